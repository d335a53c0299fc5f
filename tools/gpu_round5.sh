#!/bin/bash
# Validation pass of HEAD: smoke, GPU parity suite, headline bench (AUTO -> FP4 form), reference arm,
# ncu launch list of the same bench command, configs sweep with the current default kernels.
#   gpurun --timeout 1700 -- 'bash tools/gpu_round5.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 python tools/bench_configs.py --quick > gpurun_out/configs_quick.jsonl 2> gpurun_out/configs_quick.err; echo "configs rc=$?"
tail -n 3 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench_n1.err gpurun_out/configs_quick.err
cut -c1-2500 gpurun_out/bench_n1.json; cat gpurun_out/bench_ref.json
