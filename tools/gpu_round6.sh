#!/bin/bash
# Schedule knobs of the UMMA kernel: parity suite + shape sweeps with each knob off/on.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_gpu.log
for knob in stream_k prefill; do
  timeout 600 python tools/shape_sweep.py --knob $knob > gpurun_out/sweep_$knob.jsonl 2> gpurun_out/sweep_$knob.err; echo "sweep $knob rc=$?"; tail -n 5 gpurun_out/sweep_$knob.err
done
