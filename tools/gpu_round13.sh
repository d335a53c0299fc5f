#!/bin/bash
# Full-size configuration pass: same-box A/B of HEAD against the round's previous library on C3 itself, the whole
# C5 grid on one GPU, C1 / C4 at full size, C2 at 10 000 rows (the two heaviest levels at 2 000 rows).
#   gpurun --timeout 1500 -- 'bash tools/gpu_round13.sh'
set -u
mkdir -p gpurun_out
timeout -s KILL 200 python tools/ab_libs.py base,head 200000:131072 > gpurun_out/ab_c3_full.jsonl 2> gpurun_out/ab_c3_full.err; echo "ab c3 rc=$?"; cat gpurun_out/ab_c3_full.jsonl
timeout -s KILL 500 python tools/bench_configs.py c5grid > gpurun_out/c5_grid.jsonl 2> gpurun_out/c5_grid.err; echo "c5grid rc=$?"; cut -c1-330 gpurun_out/c5_grid.jsonl | tail -n 26; tail -n 3 gpurun_out/c5_grid.err
timeout -s KILL 400 python tools/bench_configs.py c1 c4 > gpurun_out/configs_c1_c4_full.jsonl 2> gpurun_out/configs_c1_c4_full.err; echo "c1c4 rc=$?"; cut -c1-420 gpurun_out/configs_c1_c4_full.jsonl | tail -n 14; tail -n 3 gpurun_out/configs_c1_c4_full.err
timeout -s KILL 500 python tools/bench_configs.py c2 --c2-heavy-rows=2000 > gpurun_out/configs_c2_full.jsonl 2> gpurun_out/configs_c2_full.err; echo "c2 rc=$?"; cut -c1-700 gpurun_out/configs_c2_full.jsonl | tail -n 13; tail -n 3 gpurun_out/configs_c2_full.err
