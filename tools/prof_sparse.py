#!/usr/bin/env python
"""Driver for timing / ncu of the STORM_t merge-probe kernel (sparse_pairs_kernel).

    python tools/prof_sparse.py <rows> <bits> <draws> [reps]        (JSON line on stdout)

Builds a STORM_t with `rows` rows of `draws` uniform positions (the benchmark.cpp:563-581 recipe, oracle
generator), forces the sparse route and times STORM_pairw_intersect_cardinality_blocked(s, 0); the total is
checked against the column-count closed form."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stormbitmaps_b200 as sb
from oracle import oracle as O

rows, bits, draws = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
orc = O.Oracle()
pos = [orc.gen_row_positions(2, i, draws, bits) for i in range(rows)]
cnt = np.zeros(bits, dtype=np.int64)
for p in pos:
    cnt[p] += 1
exact = int((cnt * (cnt - 1) // 2).sum())
nnz = int(sum(len(p) for p in pos))
with sb.Storm() as s:
    for p in pos:
        s.add(p)
    sb.set_storm_route("sparse")
    best, got = 1e30, None
    for r in range(reps + 1):
        t0 = time.perf_counter()
        got = s.pairw_intersect_cardinality_blocked(0)
        if r:
            best = min(best, time.perf_counter() - t0)
    W = (bits + 63) // 64
    pairs = rows * (rows - 1) / 2
    print(json.dumps({"rows": rows, "bits": bits, "draws": draws, "nnz": nnz, "route": s.last_route(), "seconds": best,
                      "pairs_per_s": pairs / best, "bitmap_space_wp_per_s": pairs * W / best,
                      "list_elements_per_s": pairs * 2 * nnz / rows / best, "match": got == exact}))
