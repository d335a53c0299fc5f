#!/usr/bin/env python
"""Dense tile kernel over a grid of shapes, with the schedule knobs A/B'd: JSON lines (run on the GPU box).

    python tools/shape_sweep.py [--knob stream_k|wave_sync|chain|none] [rows:bits ...]

Per shape and knob value: best-of-5 CUDA-event time of STORM_b200_pairw_device (rows resident, AUTO kernel),
the total checked against the column-count closed form, wp/s and the fraction of the measured mxf4 pipe.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stormbitmaps_b200 as sb
sb.load()

args = sys.argv[1:]
knob = "stream_k"
if "--knob" in args:
    k = args.index("--knob"); knob = args[k + 1]; del args[k:k + 2]
shapes = [tuple(int(x) for x in a.split(":")) for a in args] or [
    (300, 65536), (1500, 524288), (2000, 65536), (10000, 65536), (10000, 524288), (16384, 4096), (32768, 4096),
    (65536, 4096), (16384, 16384), (16384, 65536), (30000, 131072), (5000, 1048576)]
SETTERS = {"stream_k": getattr(sb, "set_umma_stream_k", None), "wave_sync": sb.set_umma_wave_sync,
           "chain": getattr(sb, "set_umma_chain", None), "none": None}
setter = SETTERS[knob]


def closed_form(rows_t, W):
    counts = torch.zeros((64, W), dtype=torch.int64, device=rows_t.device)
    for r0 in range(0, rows_t.shape[0], 8192):
        blk = rows_t[r0:r0 + 8192, :W]
        for b in range(64):
            counts[b] += ((blk >> b) & 1).sum(dim=0, dtype=torch.int64)
    return int((counts * (counts - 1) // 2).sum().item())


def timed(rows, W, reps=5):
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    sb.pairw_device(rows, n_words=W, total=total)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e30
    for _ in range(reps):
        total.zero_()
        ev[0].record()
        sb.pairw_device(rows, n_words=W, total=total)
        ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    return best, int(total.item())


peak = sb.microbench(7)[0] / 1e12
print(json.dumps({"fp4_peak_tops": peak, "knob": knob}), flush=True)
for (n, M) in shapes:
    rows, W = sb.alloc_rows(n, M)
    sb.synth_geno_device(rows, M, 1)
    torch.cuda.synchronize()
    exact = closed_form(rows, W)
    wp = n * (n - 1) / 2 * W
    rec = {"rows": n, "bits": M, "kernel": sb.resolved_kernel_name("auto", W)}
    for val in ((0, 1) if setter else (None,)):
        if setter:
            prev = setter(val)
        ms, tot = timed(rows, W)
        if setter:
            setter(prev)
        tag = f"{knob}{val}" if setter else "default"
        rec[tag] = {"ms": round(ms, 5), "wp_per_s": wp / ms * 1e3, "frac_of_fp4_peak": round(wp * 128 / ms * 1e3 / 1e12 / peak, 4),
                    "match": tot == exact}
    print(json.dumps(rec), flush=True)
    del rows
