#!/usr/bin/env python
"""FP4 (kind::mxf4) tile kernel vs the kind::i8 one: exactness on seeded shapes, then throughput.
JSON lines on stdout (run on the GPU box)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stormbitmaps_b200 as sb

sb.load()

def emit(**kw):
    print(json.dumps(kw), flush=True)

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

# ---- exactness: totals and per-pair rectangles, ragged words / rows, dense and sparse rows ----
ok_all = True
for (n, M, draws, seed) in [(300, 8192, 3000, 1), (777, 4160, 4000, 2), (1000, 64 * 67, 2000, 3), (513, 64 * 5, 200, 4),
                            (2500, 65536, 32768, 5), (260, 131072, 131072 * 3, 6), (1029, 1 << 20, 1 << 19, 7), (300, 64, 40, 8)]:
    rows, W = sb.alloc_rows(n, M)
    sb.synth_uniform_device(rows, M, draws, seed)
    t_i8 = int(sb.pairw_device(rows, n_words=W, kernel="umma").item())
    t_fp4 = int(sb.pairw_device(rows, n_words=W, kernel="fp4").item())
    i1, j1 = min(n, 300), min(n, 290)
    c_i8, _ = sb.pairw_rect_device(rows, 0, i1, 3, j1, n_words=W, kernel="umma")
    c_fp4, s_fp4 = sb.pairw_rect_device(rows, 0, i1, 3, j1, n_words=W, kernel="fp4")
    same = bool((c_i8 == c_fp4).all().item())
    ok = (t_i8 == t_fp4) and same and int(s_fp4.item()) == int(c_i8.sum().item())
    ok_all &= ok
    emit(check="exact", rows=n, bits=M, draws=draws, total_i8=t_i8, total_fp4=t_fp4, rect_equal=same, max_count=int(c_i8.max().item()), ok=ok)
# all-ones rows: every count = M (largest accumulators)
for (n, M) in [(300, 131072), (260, 1 << 22)]:
    rows, W = sb.alloc_rows(n, M)
    rows[:, :W] = -1
    t_i8 = int(sb.pairw_device(rows, n_words=W, kernel="umma").item())
    t_fp4 = int(sb.pairw_device(rows, n_words=W, kernel="fp4").item())
    ok = t_i8 == t_fp4 == n * (n - 1) // 2 * M
    ok_all &= ok
    emit(check="all_ones", rows=n, bits=M, total_i8=t_i8, total_fp4=t_fp4, ok=ok)
emit(check="summary", all_exact=ok_all)

# ---- throughput ----
for (n, M) in [(30000, 131072), (32768, 4096), (32768, 16384), (32768, 65536), (10000, 65536), (200000, 131072)]:
    rows, W = sb.alloc_rows(n, M)
    sb.synth_geno_device(rows, M, 1)
    torch.cuda.synchronize()
    wp = n * (n - 1) / 2 * W
    ms8, t8 = timed(rows, W, "umma")
    ms4, t4 = timed(rows, W, "fp4")
    emit(check="speed", rows=n, bits=M, i8_ms=ms8, fp4_ms=ms4, i8_wp_per_s=wp / ms8 * 1e3, fp4_wp_per_s=wp / ms4 * 1e3,
         fp4_tops=wp * 128 / ms4 * 1e3 / 1e12, speedup=ms8 / ms4, totals_equal=t8 == t4)
    del rows
