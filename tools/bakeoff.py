#!/usr/bin/env python
"""The five formulations of the dense path on one shape (default 30 000 x 131 072, genotype-like rows), same box, same
matrix: CUDA cores (LOP3 + POPC; carry-save), mma.sync .b1 AND + POPC (emulated on sm_100a), tcgen05 kind::i8 on
bits unpacked to bytes, tcgen05 kind::mxf4 on bits unpacked to E2M1 nibbles.  Best of 3 CUDA-event times of
STORM_b200_pairw_device, every total checked against the column-count closed form.  JSON lines (run on the GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stormbitmaps_b200 as sb
sb.load()
n, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (30000, 131072)
rows, W = sb.alloc_rows(n, M)
sb.synth_geno_device(rows, M, 1)
torch.cuda.synchronize()
counts = torch.zeros((64, W), dtype=torch.int64, device="cuda")
for r0 in range(0, n, 8192):
    blk = rows[r0:r0 + 8192, :W]
    for b in range(64):
        counts[b] += ((blk >> b) & 1).sum(dim=0, dtype=torch.int64)
exact = int((counts * (counts - 1) // 2).sum().item())
wp = n * (n - 1) / 2 * W
total = torch.zeros(1, dtype=torch.int64, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for kernel in ("popc", "csa", "b1", "umma", "fp4"):
    best = 1e30
    for rep in range(4):
        total.zero_()
        ev[0].record()
        sb.pairw_device(rows, n_words=W, kernel=kernel, total=total)
        ev[1].record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, ev[0].elapsed_time(ev[1]))
    print(json.dumps({"rows": n, "bits": M, "kernel": kernel, "ms": round(best, 3), "wp_per_s": wp / best * 1e3,
                      "match": int(total.item()) == exact}), flush=True)
