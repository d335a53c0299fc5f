#!/bin/bash
# One GPU-box pass: parity tests, headline bench, reference arm, ncu launch list and one full capture.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_umma -c 1 -o gpurun_out/umma_c3_full -f \
    python tools/prof_driver.py umma 200000 131072 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json gpurun_out/bench_ref.json
