#!/bin/bash
# ncu captures: small-K shape (where does a tile's time go) and the C3 default (DRAM traffic for traffic.json).
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_umma -s 2 -c 1 -o gpurun_out/fp4_65536x4096_full -f \
    python tools/prof_driver.py fp4 65536 4096 3 > gpurun_out/ncu_small.log 2>&1; echo "ncu small rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_umma -s 1 -c 1 -o gpurun_out/fp4_c3_full_v3 -f \
    python tools/prof_driver.py fp4 200000 131072 2 > gpurun_out/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
ls -la gpurun_out/*.ncu-rep
