#!/usr/bin/env python
"""FP4 tile kernel: one vs two expander warps per 32 rows (STORM_b200_set_umma_variant bit 3), exactness
against the int8 form and throughput on a few shapes.  JSON lines (run on the GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stormbitmaps_b200 as sb
sb.load()

def timed(rows, W, kernel, reps=3):
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    sb.pairw_device(rows, n_words=W, kernel=kernel, total=total)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e30
    for _ in range(reps):
        total.zero_()
        ev[0].record()
        sb.pairw_device(rows, n_words=W, kernel=kernel, total=total)
        ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    return best, int(total.item())

peak = sb.microbench(7)[0] / 1e12
peak8 = sb.microbench(5)[0] / 1e12
for ws in (1, 0):
    sb.set_umma_wave_sync(bool(ws))
    for (n, M) in [(30000, 131072), (32768, 4096), (10000, 65536), (200000, 131072)]:
        if n == 200000 and ws == 0:
            continue
        rows, W = sb.alloc_rows(n, M)
        sb.synth_geno_device(rows, M, 1)
        torch.cuda.synchronize()
        wp = n * (n - 1) / 2 * W
        ms, _ = timed(rows, W, "umma")
        print(json.dumps({"kernel": "i8", "wave_sync": ws, "rows": n, "bits": M, "ms": ms, "wp_per_s": wp / ms * 1e3,
                          "frac_of_i8_peak": wp * 128 / ms * 1e3 / 1e12 / peak8}), flush=True)
        sb.set_umma_variant(3)
        ms, _ = timed(rows, W, "fp4")
        print(json.dumps({"kernel": "fp4", "wide": 0, "wave_sync": ws, "rows": n, "bits": M, "ms": ms, "wp_per_s": wp / ms * 1e3,
                          "frac_of_fp4_peak": wp * 128 / ms * 1e3 / 1e12 / peak}), flush=True)
        del rows
sb.set_umma_wave_sync(True)
for wide in (0, 1):
    sb.set_umma_variant(3 | (8 if wide else 0))
    ok = True
    for (n, M, draws, seed) in [(777, 4160, 4000, 2), (513, 320, 200, 4), (2500, 65536, 32768, 5), (300, 64, 40, 8), (1029, 1 << 20, 1 << 19, 7)]:
        rows, W = sb.alloc_rows(n, M)
        sb.synth_uniform_device(rows, M, draws, seed)
        a = int(sb.pairw_device(rows, n_words=W, kernel="umma").item())
        b = int(sb.pairw_device(rows, n_words=W, kernel="fp4").item())
        c8, _ = sb.pairw_rect_device(rows, 0, min(n, 300), 3, min(n, 290), n_words=W, kernel="umma")
        c4, _ = sb.pairw_rect_device(rows, 0, min(n, 300), 3, min(n, 290), n_words=W, kernel="fp4")
        ok &= (a == b) and bool((c8 == c4).all().item())
    print(json.dumps({"wide": wide, "exact_vs_i8": ok}), flush=True)
    for (n, M) in [(30000, 131072), (32768, 4096), (32768, 16384), (10000, 65536), (200000, 131072)]:
        rows, W = sb.alloc_rows(n, M)
        sb.synth_geno_device(rows, M, 1)
        torch.cuda.synchronize()
        wp = n * (n - 1) / 2 * W
        ms, _ = timed(rows, W, "fp4")
        print(json.dumps({"wide": wide, "rows": n, "bits": M, "ms": ms, "wp_per_s": wp / ms * 1e3, "tops": wp * 128 / ms * 1e3 / 1e12,
                          "frac_of_fp4_peak": wp * 128 / ms * 1e3 / 1e12 / peak}), flush=True)
        del rows
