#!/bin/bash
# Accumulator chaining (one drain per run of interior tiles): parity, same-box A/B against HEAD~ (lib_base) and
# the knob A/B over the shape grid, then the whole GPU suite.
#   tools/build_variant.sh HEAD base; tools/build_variant.sh WORK chain
#   gpurun --timeout 1200 -- 'bash tools/gpu_round11.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout -s KILL 300 python -m pytest tests/test_parity_gpu.py -x -q -k "chaining or stream_k or wave_sync" > gpurun_out/pytest_chain.log 2>&1; echo "pytest chain rc=$?"
tail -n 15 gpurun_out/pytest_chain.log
timeout -s KILL 240 python tools/ab_libs.py base,chain 16384:4096 65536:4096 32768:16384 10000:65536 30000:131072 > gpurun_out/ab_chain.jsonl 2> gpurun_out/ab_chain.err; echo "ab rc=$?"
cat gpurun_out/ab_chain.jsonl; tail -n 3 gpurun_out/ab_chain.err
timeout -s KILL 300 python tools/shape_sweep.py --knob chain > gpurun_out/sweep_chain.jsonl 2> gpurun_out/sweep_chain.err; echo "sweep rc=$?"
cut -c1-400 gpurun_out/sweep_chain.jsonl; tail -n 3 gpurun_out/sweep_chain.err
timeout -s KILL 700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/pytest_gpu.log
