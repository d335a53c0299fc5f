#!/bin/bash
# Validation of HEAD after accumulator chaining / stream-K under wave sync / vector stores: smoke, GPU suite,
# both bench arms, ncu launch list of the bench command, ncu --set full of the default kernel on C3 and on 65536 x 4096.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round12.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout -s KILL 400 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"; cut -c1-500 gpurun_out/bench_ref_n1.json
timeout -s KILL 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/bench_n1.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_v4.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:dense_umma -s 2 -c 1 -o gpurun_out/fp4_65536x4096_full_v2 -f \
    python tools/prof_driver.py fp4 65536 4096 3 > gpurun_out/ncu_small.log 2>&1; echo "ncu small rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:dense_umma -s 1 -c 1 -o gpurun_out/fp4_c3_full_v4 -f \
    python tools/prof_driver.py fp4 200000 131072 2 > gpurun_out/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
ls -la gpurun_out/*.ncu-rep
