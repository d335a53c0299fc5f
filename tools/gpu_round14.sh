#!/bin/bash
# Round-end validation of HEAD: smoke, whole GPU suite, both bench arms, knob sweep of chaining over the shape grid.
#   gpurun --timeout 1200 -- 'bash tools/gpu_round14.sh'
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout -s KILL 400 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_n1.json
timeout -s KILL 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-2400 gpurun_out/bench_n1.json
timeout -s KILL 300 python tools/shape_sweep.py --knob chain > gpurun_out/sweep_chain_v2.jsonl 2> gpurun_out/sweep_chain_v2.err; echo "sweep rc=$?"
