#!/usr/bin/env python
"""Smoke-sized run of every kernel family for compute-sanitizer ("race detection"):

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Small shapes (a few tiles, ragged edges, a K tail) through the tensor kernels <2, FP4> and <2, i8> (totals with
stream-K and chaining, per-pair rectangles), the CUDA-core kernels, the mma.sync b1 kernel, the contiguous-model list
routes (probe + stream kernels) and the three STORM_t kernels -- each checked against the oracle, so a sanitizer run
is also a correctness run.  Prints one line per step; exit code 0 only if every value matched."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import stormbitmaps_b200 as sb  # noqa: E402
from oracle import oracle as O  # noqa: E402

sb.load()
orc = O.Oracle()
ok = True


def check(name, got, want):
    global ok
    good = got == want if not isinstance(got, np.ndarray) else bool((got == want).all())
    ok = ok and good
    print(f"{'ok ' if good else 'BAD'} {name}", flush=True)


def dev_rows(vals):
    n, w = vals.shape
    rows, _ = sb.alloc_rows(n, w * 64)
    rows[:, :w] = torch.from_numpy(vals.view(np.int64)).cuda()
    return rows


for (M, N, draws) in [(2048 + 64, 600, 700), (192, 300, 60)]:
    vals = orc.gen_dense_uniform(3, N, draws, M)
    W = vals.shape[1]
    rows = dev_rows(vals)
    exact = orc.wrapper_diag(vals)
    for kernel in ("fp4", "umma", "popc", "csa", "b1"):
        check(f"total {kernel} {N}x{M}", int(sb.pairw_device(rows, n_words=W, kernel=kernel).item()), exact)
    for kernel in ("fp4", "umma"):
        got, tot = sb.pairw_rect_device(rows, 10, 290, 37, N - 1, n_words=W, kernel=kernel)
        check(f"rect {kernel} {N}x{M}", got.cpu().numpy().view(np.uint32), orc.rect_counts(vals, 10, 290, 37, N - 1))
    parts = sum(int(sb.pairw_device(rows, n_words=W, shard=k, n_shards=3, kernel="fp4").item()) for k in range(3))
    check(f"shards fp4 {N}x{M}", parts, exact)

# chained accumulators + stream-K tail: many small interior tiles
vals = orc.gen_dense_uniform(5, 2600, 100, 256)
rows = dev_rows(vals)
check("chain fp4 2600x256", int(sb.pairw_device(rows, n_words=4, kernel="fp4").item()), orc.wrapper_diag(vals))
check("chain umma 2600x256", int(sb.pairw_device(rows, n_words=4, kernel="umma").item()), orc.wrapper_diag(vals))

# contiguous model: list routes (probe kernel, stream kernel), replicas on one device
M = 65536
rows_p = [orc.gen_row_positions(31, i, [1, 5, 60, 150, 199, 3][i % 6], M) for i in range(260)]
exact = orc.wrapper_diag(O.positions_to_dense(rows_p, M))
mixed = [orc.gen_row_positions(32, i, [5, 150, 4000, 30000][i % 4], M) for i in range(260)]
exact_mixed = orc.wrapper_diag(O.positions_to_dense(mixed, M))
for ids in ((), (0, 0)):
    sb.set_device_list(ids)
    with sb.StormContiguous(M) as c, sb.StormContiguous(M) as d:
        for p in rows_p:
            c.add(p)
        for p in mixed:
            d.add(p)
        for route in ("stream", "probe", "tile"):
            prev = sb.set_contig_list_route(route)
            check(f"contig list {route} devices={ids or 'default'}", c.pairw_intersect_cardinality_list(), exact)
            check(f"contig list mixed {route} devices={ids or 'default'}", d.pairw_intersect_cardinality_list(), exact_mixed)
            sb.set_contig_list_route(prev)
        check(f"contig blocked devices={ids or 'default'}", d.pairw_intersect_cardinality_blocked(31), exact_mixed)
    check(f"wrapper devices={ids or 'default'}", sb.wrapper_diag(O.positions_to_dense(mixed, M)), exact_mixed)
sb.set_device_list(())

# STORM_t: block, flat and stream kernels + the densified route
M = 3 * 65536 + 1000
draws = [1, 5, 40, 64, 65, 300, 3000, 9000, 60000, 0]
srows = [orc.gen_row_positions(55, i, draws[i % len(draws)], M) for i in range(120)]
sexact = orc.wrapper_diag(O.positions_to_dense(srows, M))
lrows = [orc.gen_row_positions(56, i, [0, 1, 3, 17, 64, 200, 1000][i % 7], M) for i in range(150)]
lexact = orc.wrapper_diag(O.positions_to_dense(lrows, M))
for route in ("sparse", "dense", "split"):
    prev = sb.set_storm_route(route)
    with sb.Storm() as s:
        for p in srows:
            s.add(p)
        check(f"storm_t mixed route={route}", s.pairw_intersect_cardinality(), sexact)
        if route == "split":
            check("storm_t split route shards", sum(s.pairw_shard(k, 3) for k in range(3)), sexact)
            sb.set_storm_route(prev)
            continue
    for flat in ("stream", "flat", "block"):
        was = sb.set_sparse_flat(flat)
        with sb.Storm() as s:
            for p in lrows:
                s.add(p)
            check(f"storm_t lists route={route} kernel={flat}", s.pairw_intersect_cardinality(), lexact)
        sb.set_sparse_flat(was)
    sb.set_storm_route(prev)

# STORM_t on replicas, banded dense route, rectangles through densified rows
for ids in ((0, 0),):
    sb.set_device_list(ids)
    for route in ("sparse", "dense", "split"):
        prev = sb.set_storm_route(route)
        with sb.Storm() as s:
            for p in srows:
                s.add(p)
            check(f"storm_t replicas route={route}", s.pairw_intersect_cardinality(), sexact)
        sb.set_storm_route(prev)
    sb.set_device_list(())
prev = sb.set_storm_route("dense")
was = sb.set_storm_band_rows(256)
with sb.Storm() as s:
    for p in (srows * 6)[:600]:
        s.add(p)
    vals6 = O.positions_to_dense((srows * 6)[:600], M)
    check("storm_t banded dense", s.pairw_intersect_cardinality(), orc.wrapper_diag(vals6))
    check("storm_t rect via densified rows", s.pairw_rect(7, 130, 60, 420), orc.rect_counts(vals6, 7, 130, 60, 420))
sb.set_storm_band_rows(was)
sb.set_storm_route(prev)

torch.cuda.synchronize()
print("sanitize_smoke:", "ALL OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
