#!/usr/bin/env python
"""Per-pair output path: STORM_b200_pairw_rect_device writing uint32 counts of an
interior rectangle, timed with CUDA events; JSON lines (run on the GPU box).

    python tools/rect_bench.py [rows:bits:side ...]

Reports pairs/s, the output bandwidth (4 B per pair) and the same rectangle as a total-only query."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stormbitmaps_b200 as sb
L = sb.load()
from stormbitmaps_b200 import _lib

variants = [None]                                 # --variants=51,35: STORM_b200_set_umma_variant values to time the counts with (35 = without the TMA-store drain)
for a in list(sys.argv[1:]):
    if a.startswith("--variants="):
        variants = [int(x) for x in a.split("=")[1].split(",")]
        sys.argv.remove(a)
shapes = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(32768, 4096, 16384), (32768, 16384, 16384), (32768, 65536, 16384), (24576, 131072, 12288)]
for (n, M, side) in shapes:
    rows, W = sb.alloc_rows(n, M)
    sb.synth_geno_device(rows, M, 3)
    stride = rows.shape[1]
    out = torch.zeros((side, side), dtype=torch.int32, device="cuda")
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    rec = {"rows": n, "bits": M, "rect": [side, side]}
    runs = [("counts" if v is None else f"counts_variant{v}", out.data_ptr(), v) for v in variants] + [("total_only", None, None)]
    for tag, optr, var in runs:
        prev_var = sb.set_umma_variant(var) if var is not None else None
        best = 1e30
        for rep in range(5):
            total.zero_()
            ev[0].record()
            _lib.check(L.STORM_b200_pairw_rect_device(rows.data_ptr(), n, W, stride, 0, side, n - side, n, 0, 0, optr, side,
                                                      total.data_ptr(), None), "rect")
            ev[1].record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, ev[0].elapsed_time(ev[1]))
        if prev_var is not None:
            sb.set_umma_variant(prev_var)
        pairs = side * side
        rec[tag] = {"ms": round(best, 4), "pairs_per_s": pairs / best * 1e3, "wp_per_s": pairs * W / best * 1e3,
                    "out_GBps": (pairs * 4 / best * 1e3 / 1e9) if optr else 0.0, "total": int(total.item())}
    first = "counts" if variants[0] is None else f"counts_variant{variants[0]}"
    rec["counts_sum_matches_total"] = int(out.sum(dtype=torch.int64).item()) == rec[first]["total"] == rec["total_only"]["total"]
    print(json.dumps(rec), flush=True)
    del rows, out
