#!/usr/bin/env python
"""Tiny driver for ncu: builds a resident matrix and runs one kernel variant a few times.

    python tools/prof_driver.py <kernel> <rows> <bits> [reps] [umma_cg]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stormbitmaps_b200 as sb
kernel = sys.argv[1]
rows, bits = int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
if len(sys.argv) > 5:
    sb.set_umma_cta_group(int(sys.argv[5]))
t, W = sb.alloc_rows(rows, bits)
sb.synth_geno_device(t, bits, 1)
total = torch.zeros(1, dtype=torch.int64, device="cuda")
for _ in range(reps):
    total.zero_()
    sb.pairw_device(t, n_words=W, kernel=kernel, total=total)
torch.cuda.synchronize()
print(kernel, rows, bits, int(total.item()))
