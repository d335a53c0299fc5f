#!/usr/bin/env python
"""UMMA kernel throughput against row width (the per-tile epilogue is the overhead that shows at small W)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stormbitmaps_b200 as sb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
for M in (2048, 4096, 8192, 16384, 32768, 65536, 131072):
    W = M // 64
    t, _ = sb.alloc_rows(N, M)
    sb.synth_uniform_device(t, M, M // 2, 9)
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    ref = int(sb.pairw_device(t, n_words=W, kernel="csa").item()) if M <= 8192 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for rep in range(4):
        total.zero_(); e0.record()
        sb.pairw_device(t, n_words=W, kernel="umma", total=total)
        e1.record(); torch.cuda.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1) * 1e-3)
    wp = N * (N - 1) / 2 * W
    print(json.dumps({"rows": N, "bits": M, "ms": best * 1e3, "wp_per_s": wp / best, "tops": wp * 128 / best / 1e12,
                      "match_csa": None if ref is None else ref == int(total.item())}), flush=True)
    del t
