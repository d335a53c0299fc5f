#!/bin/bash
# Round-end validation of HEAD after the sparse-route kernels: smoke, whole GPU suite, default bench.
#   gpurun --timeout 900 -- 'bash tools/gpu_round15.sh'
set -u
mkdir -p gpurun_out
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout -s KILL 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout -s KILL 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_n1.json
