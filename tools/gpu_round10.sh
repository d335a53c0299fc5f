#!/bin/bash
# Round-1 validation of HEAD: smoke, GPU parity suite, default bench (both arms), sparse-kernel timings.
#   gpurun --timeout 1300 -- 'bash tools/gpu_round10.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref_n1.json
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/bench_n1.json
rm -f gpurun_out/sparse_timing.jsonl
for cfg in "10000 524288 104" "10000 524288 5242" "3000 1048576 10486"; do
  timeout 300 python tools/prof_sparse.py $cfg >> gpurun_out/sparse_timing.jsonl 2>> gpurun_out/sparse_timing.err; echo "sparse $cfg rc=$?"
done
cat gpurun_out/sparse_timing.jsonl
