#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/fp4_tune.py > gpurun_out/fp4_tune.jsonl 2> gpurun_out/fp4_tune.err; echo "tune rc=$?"; cat gpurun_out/fp4_tune.jsonl; tail -n 5 gpurun_out/fp4_tune.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_umma -c 1 -o gpurun_out/fp4w_c3_full -f \
    python tools/prof_driver.py fp4 200000 131072 1 > gpurun_out/ncu_fp4w.log 2>&1
