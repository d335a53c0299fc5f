#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/fp4_tune.py > gpurun_out/fp4_tune2.jsonl 2> gpurun_out/fp4_tune2.err; echo "tune rc=$?"; cat gpurun_out/fp4_tune2.jsonl; tail -n 5 gpurun_out/fp4_tune2.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err; echo "bench rc=$?"; cat gpurun_out/bench_auto.json | cut -c1-1500; tail -n 3 gpurun_out/bench_auto.err
