#!/usr/bin/env python
"""A/B of the L2 eviction hints on the packed-row loads (STORM_b200_set_umma_l2_hints 0 / 1 / 2) on one box: C3 (or
rows:bits given), resident matrix, interleaved rounds, CUDA-event times; JSON lines.  Run once more under
`ncu --metrics dram__bytes_read.sum` for the DRAM traffic of each mode (launch order: mode 0, 1, 2, 0, 1, 2, ...)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stormbitmaps_b200 as sb
L = sb.load()
n, M = (int(x) for x in sys.argv[1].split(":")) if len(sys.argv) > 1 else (200000, 131072)
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows, W = sb.alloc_rows(n, M)
sb.synth_geno_device(rows, M, 20260117)
torch.cuda.synchronize()
total = torch.zeros(1, dtype=torch.int64, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
res = {0: [], 1: [], 2: []}
tot = {}
for r in range(rounds + 1):
    for mode in (0, 1, 2):
        L.STORM_b200_set_umma_l2_hints(mode)
        total.zero_()
        ev[0].record()
        sb.pairw_device(rows, n_words=W, total=total)
        ev[1].record()
        torch.cuda.synchronize()
        if r:
            res[mode].append(ev[0].elapsed_time(ev[1]))
        tot[mode] = int(total.item())
wp = n * (n - 1) / 2 * W
for mode in (0, 1, 2):
    best = min(res[mode])
    print(json.dumps({"rows": n, "bits": M, "l2_hints": mode, "ms": res[mode], "best_ms": best, "wp_per_s": wp / best * 1e3,
                      "totals_agree": tot[mode] == tot[0]}), flush=True)
