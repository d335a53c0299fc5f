#!/bin/bash
# Round-1 re-entry pass: deferred-epilogue A/B (same box), validation of the working tree, sparse-kernel
# timing + ncu capture, full-size C1 / C4 / C5 cells.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round9.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python tools/ab_libs.py base,d0,d1 16384:4096 65536:4096 16384:16384 65536:16384 10000:65536 30000:131072 2000:65536 \
    > gpurun_out/ab_defer.jsonl 2> gpurun_out/ab_defer.err; echo "ab rc=$?"; cat gpurun_out/ab_defer.jsonl | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_n1.json
for cfg in "10000 524288 104" "10000 524288 5" "3000 524288 5242" "3000 1048576 10486"; do
  timeout 300 python tools/prof_sparse.py $cfg >> gpurun_out/sparse_timing.jsonl 2>> gpurun_out/sparse_timing.err; echo "sparse $cfg rc=$?"
done
cat gpurun_out/sparse_timing.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sparse_pairs -s 1 -c 1 -o gpurun_out/sparse_c4_full -f \
    python tools/prof_sparse.py 3000 1048576 10486 1 > gpurun_out/ncu_sparse.log 2>&1; echo "ncu sparse rc=$?"
timeout 900 python tools/bench_configs.py c1 c5 > gpurun_out/configs_full.jsonl 2> gpurun_out/configs_full.err; echo "configs rc=$?"
tail -n 3 gpurun_out/configs_full.err
python - <<'P'
import json
for l in open('gpurun_out/configs_full.jsonl'):
    d = json.loads(l)
    if 'config' in d:
        print(d['config'], d.get('rows'), d.get('bits'), d.get('draws'), 'wp/s %.4g' % float(d.get('wp_per_s', 0)), d.get('match'))
P
