#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python tools/shape_sweep.py --knob flags 2000:65536 10000:65536 16384:4096 32768:4096 65536:4096 16384:16384 16384:65536 30000:131072 > gpurun_out/sweep_flags.jsonl 2> gpurun_out/sweep_flags.err; echo "sweep rc=$?"; tail -n 5 gpurun_out/sweep_flags.err
