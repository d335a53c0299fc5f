#!/usr/bin/env python
"""Throughput + exactness of the BASELINE.json configurations other than the bench.py headline (C3).

    python tools/bench_configs.py [c1] [c2] [c4] [c5] [c5grid] [--quick] [--c2-heavy-rows=N]      (JSON lines on stdout)

C1  benchmark 65536 10000 (benchmark.cpp:693-698 density levels): STORM_contiguous_t through the
    storm.h API, rows ingested with STORM_contig_add / STORM_b200_contig_add_bulk.
C2  STORM_t 10,000 x 524,288 density sweep (benchmark.cpp:511, README.md:65-80) + one mixed level.
C4  100,000 x 1,048,576 at ~1 % (10,486 draws per row): STORM_t (all list blocks) and the dense model.
C5  scaling sweep cells on one GPU, device-resident rows (c5grid: all 25 cells of the N x M grid).
--c2-heavy-rows=N: row count of the C2 levels above 60,000 draws per row (host-side generation and the
    host closed form of those levels take minutes at 10,000 rows; they take the same code path as 52,428).

Every total is checked against the column-count closed form sum_k C(c_k, 2) (an O(N*W) identity
that shares no code with the kernels); the timed region is the query call with rows resident
(the reference's PERF_PRE/PERF_POST placement, benchmark.cpp:906-911).  wp/s for the sparse
model is the README's bitmap-space-equivalent unit N(N-1)/2 * ceil(M/64) / s.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import stormbitmaps_b200 as sb  # noqa: E402
from oracle import oracle as O  # noqa: E402

QUICK = "--quick" in sys.argv
which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c2", "c4", "c5"]
C2_HEAVY_ROWS = next((int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--c2-heavy-rows=")), None)
orc = O.Oracle()


def emit(**kw):
    print(json.dumps(kw), flush=True)


def closed_form(positions_per_row, M):
    cnt = np.zeros(M, dtype=np.int64)
    for p in positions_per_row:
        cnt[np.unique(p)] += 1
    return int((cnt * (cnt - 1) // 2).sum())


def closed_form_device(rows_t, W):
    counts = torch.zeros((64, W), dtype=torch.int64, device=rows_t.device)
    for r0 in range(0, rows_t.shape[0], 8192):
        blk = rows_t[r0:r0 + 8192, :W]
        for b in range(64):
            counts[b] += ((blk >> b) & 1).sum(dim=0, dtype=torch.int64)
    return int((counts * (counts - 1) // 2).sum().item())


def best_of(fn, reps=3):
    """Best wall-clock time of fn().  The GPU drops to idle clocks while the host prepares the next level, so
    short calls are repeated until 0.25 s have passed (at least `reps` times) before the best is taken."""
    fn()                                   # warm-up: uploads / mirrors become resident
    best, val, n, t_start = 1e30, None, 0, time.perf_counter()
    while n < reps or (time.perf_counter() - t_start < 0.25 and n < 200):
        t0 = time.perf_counter()
        val = fn()
        best = min(best, time.perf_counter() - t0)
        n += 1
    return best, val


def gen_rows(seed, N, draws, M):
    if isinstance(draws, int):
        return [orc.gen_row_positions(seed, i, draws, M) for i in range(N)]
    return [orc.gen_row_positions(seed, i, int(d), M) for i, d in enumerate(draws)]


def run_c1():
    M, N = 65536, 2000 if QUICK else 10000
    W = M // 64
    for draws in (32768, 16384, 6553, 2621, 1310, 655, 262, 65, 13, 5, 1):
        rows = gen_rows(1, N, draws, M)
        exact = closed_form(rows, M)
        with sb.StormContiguous(M) as c:
            off = np.zeros(N + 1, dtype=np.uint64)
            off[1:] = np.cumsum([len(r) for r in rows])
            t0 = time.perf_counter()
            c.add_bulk(np.concatenate(rows), off)
            ingest = time.perf_counter() - t0
            dt, got = best_of(lambda: c.pairw_intersect_cardinality_blocked(31))
            dt_l, got_l = best_of(lambda: c.pairw_intersect_cardinality_blocked_list(31))
            emit(config="c1", model="STORM_contiguous_t", rows=N, bits=M, draws=draws, total=got, exact=exact,
                 match=got == exact and got_l == exact, seconds=dt, wp_per_s=N * (N - 1) / 2 * W / dt,
                 list_seconds=dt_l, list_wp_per_s=N * (N - 1) / 2 * W / dt_l, ingest_seconds=ingest,
                 call="STORM_contig_pairw_intersect_cardinality_blocked(c, 31) / _blocked_list")


def run_c2():
    M, N_all = 524288, 1500 if QUICK else 10000
    W = M // 64
    levels = [262144, 131072, 52428, 20971, 10485, 5242, 2097, 524, 104, 5, 1, "mixed"]
    rng = np.random.default_rng(7)
    for draws in levels:
        N = C2_HEAVY_ROWS if (C2_HEAVY_ROWS and isinstance(draws, int) and draws > 60000) else N_all
        d = draws if draws != "mixed" else np.exp(rng.uniform(0, np.log(262144), N)).astype(np.int64)
        rows = gen_rows(2, N, d, M)
        exact = closed_form(rows, M)
        with sb.Storm() as s:
            t0 = time.perf_counter()
            for p in rows:
                s.add(p)
            build = time.perf_counter() - t0
            res = {}
            for route in ("auto", "sparse", "dense"):
                if route == "sparse" and not QUICK and isinstance(draws, int) and draws > 60000:
                    continue                # minutes on the merge/probe kernel; the route is covered by the tests
                sb.set_storm_route(route)
                dt, got = best_of(lambda: s.pairw_intersect_cardinality_blocked(0), reps=2)
                res[route] = {"seconds": dt, "wp_per_s": N * (N - 1) / 2 * W / dt, "match": got == exact,
                              "took": s.last_route()}
            sb.set_storm_route("auto")
            emit(config="c2", model="STORM_t", rows=N, bits=M, draws=draws, exact=exact, build_seconds=build,
                 serialized_size=s.serialized_size(), routes=res, match=all(r["match"] for r in res.values()),
                 call="STORM_pairw_intersect_cardinality_blocked(s, 0)", unit="bitmap-space-equivalent wp/s")


def run_c4():
    M, N, draws = 1048576, 8000 if QUICK else 100000, 10486
    W = M // 64
    # dense model, rows generated on the device (bit-identical to the oracle generator)
    rows_t, _ = sb.alloc_rows(N, M)
    sb.synth_uniform_device(rows_t, M, draws, 4)
    torch.cuda.synchronize()
    exact = closed_form_device(rows_t, W)
    total = torch.zeros(1, dtype=torch.int64, device="cuda")

    def dense():
        total.zero_()
        sb.pairw_device(rows_t, n_words=W, total=total)
        return int(total.item())
    dt, got = best_of(dense)
    emit(config="c4", model="dense rows (STORM_b200_pairw_device)", rows=N, bits=M, draws=draws, total=got, exact=exact,
         match=got == exact, seconds=dt, wp_per_s=N * (N - 1) / 2 * W / dt)
    del rows_t
    torch.cuda.empty_cache()
    # sparse model: every block is a u16 list block (655 values per block on average)
    t0 = time.perf_counter()
    with sb.Storm() as s:
        for i in range(N):
            s.add(orc.gen_row_positions(4, i, draws, M))
        build = time.perf_counter() - t0
        dt, got = best_of(lambda: s.pairw_intersect_cardinality_blocked(0), reps=2)
        emit(config="c4", model="STORM_t", rows=N, bits=M, draws=draws, total=got, exact=exact, match=got == exact,
             seconds=dt, wp_per_s=N * (N - 1) / 2 * W / dt, took=s.last_route(), build_seconds=build,
             serialized_size=s.serialized_size(), unit="bitmap-space-equivalent wp/s")


def run_c5(grid=False):
    cells = [(16384, 4096), (16384, 65536), (65536, 4096), (65536, 16384), (65536, 65536), (16384, 1048576)]
    if not QUICK:
        cells += [(131072, 65536), (262144, 16384), (65536, 262144)]
    if grid:        # N in 16k..256k x M in 4k..1M (largest cell 34.4 GB of rows)
        cells = [(N, M) for N in (16384, 32768, 65536, 131072, 262144) for M in (4096, 16384, 65536, 262144, 1048576)]
    peak = sb.microbench(7)[0] if grid else None
    for N, M in cells:
        W = M // 64
        rows_t, _ = sb.alloc_rows(N, M)
        if grid:
            sb.synth_geno_device(rows_t, M, 5)          # (the uniform generator draws M / 2 positions per row: minutes at 34 GB)
        else:
            sb.synth_uniform_device(rows_t, M, M // 2, 5)
        torch.cuda.synchronize()
        exact = closed_form_device(rows_t, W)
        total = torch.zeros(1, dtype=torch.int64, device="cuda")
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for rep in range(3):
            total.zero_()
            ev0.record()
            sb.pairw_device(rows_t, n_words=W, total=total)
            ev1.record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, ev0.elapsed_time(ev1) * 1e-3)
        got = int(total.item())
        emit(config="c5", rows=N, bits=M, total=got, exact=exact, match=got == exact, seconds=best,
             wp_per_s=N * (N - 1) / 2 * W / best, tops=N * (N - 1) / 2 * W * 128 / best / 1e12,
             **({"frac_of_mxf4_pipe": N * (N - 1) / 2 * W * 128 / best / peak, "generator": "geno"} if grid else {}))
        del rows_t
        torch.cuda.empty_cache()


if __name__ == "__main__":
    emit(device=sb.device_info(0), quick=QUICK)
    for w in which:
        {"c1": run_c1, "c2": run_c2, "c4": run_c4, "c5": run_c5, "c5grid": lambda: run_c5(grid=True)}[w]()
