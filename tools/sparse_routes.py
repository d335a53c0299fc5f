#!/usr/bin/env python
"""STORM_t whole-container query through every route on one container: row-group stream kernel, flat probe kernel, block merge/probe
kernel, densified rows + tensor kernel, and what AUTO picks.  JSON lines (run on the GPU box); the timings are what
the route cost model in sparse.cu (choose_route) is fitted to.

    python tools/sparse_routes.py [rows:bits:draws[:n_heavy:heavy_draws] ...]

With n_heavy > 0 that many rows, spread evenly, are drawn with heavy_draws values instead (rows with bitmap blocks among
light ones: what the split route is for).
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stormbitmaps_b200 as sb
from oracle import oracle as O

cases = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [
    (10000, 524288, 1), (10000, 524288, 5), (10000, 524288, 30), (10000, 524288, 104), (10000, 524288, 524),
    (10000, 524288, 2097), (3000, 1048576, 104), (3000, 1048576, 1000), (3000, 1048576, 10486), (20000, 131072, 16)]
orc = O.Oracle()


def timed(s, min_s=0.2):
    s.pairw_intersect_cardinality_blocked(0)                      # mirrors resident
    best, got, n, t_start = 1e30, None, 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t_start < min_s and n < 100):
        t0 = time.perf_counter()
        got = s.pairw_intersect_cardinality_blocked(0)
        best = min(best, time.perf_counter() - t0)
        n += 1
    return best, got


for case in cases:
    rows, bits, draws = case[:3]
    n_heavy, heavy_draws = (case[3], case[4]) if len(case) >= 5 else (0, 0)
    pos = [orc.gen_row_positions(2, i, draws, bits) for i in range(rows)]
    for k in range(n_heavy):
        r = (2 * k + 1) * rows // (2 * n_heavy)
        pos[r] = orc.gen_row_positions(3, r, heavy_draws, bits)
    cnt = np.zeros(bits, dtype=np.int64)
    for p in pos:
        cnt[p] += 1
    exact = int((cnt * (cnt - 1) // 2).sum())
    nnz = int(sum(len(p) for p in pos))
    rec = {"rows": rows, "bits": bits, "draws": draws, "n_heavy": n_heavy, "heavy_draws": heavy_draws, "avg_nnz": nnz / rows, "routes": {}}
    with sb.Storm() as s:
        for p in pos:
            s.add(p)
        for name, route, flat in (("stream", "sparse", 2), ("flat", "sparse", 1), ("block", "sparse", 0), ("split", "split", 2), ("dense", "dense", 2), ("auto", "auto", 2)):
            if n_heavy and name in ("flat", "block"):
                continue                                           # (with heavy rows "stream" already is the block kernel)
            if bits > (1 << 22) and name == "dense" and rows * rows * (bits / 64) / 2 > 3e14:
                continue
            if (name == "block" and rows * (rows - 1) / 2 * (1 + nnz / rows / 100) > 2e8) or (name == "flat" and rows * (rows - 1) / 2 * nnz / rows > 3e10):
                continue                                           # seconds on the block kernel
            sb.set_storm_route(route)
            sb.set_sparse_flat(flat)
            dt, got = timed(s)
            rec["routes"][name] = {"ms": round(dt * 1e3, 4), "took": s.last_route(), "match": got == exact}
        sb.set_storm_route("auto")
        sb.set_sparse_flat(2)
    print(json.dumps(rec), flush=True)
