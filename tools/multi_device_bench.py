#!/usr/bin/env python
"""Multi-GPU behind the C ABI: one process, one calling thread, G devices (run on a multi-GPU box: `gpurun --gpus 8`).

    python tools/multi_device_bench.py c3 [grid] [--devices 1,2,4,8]        (JSON lines on stdout)

c3    BASELINE config C3 (200 000 x 131 072, genotype-like) for G in --devices:
        resident    STORM_b200_pairw_devices: the matrix already on every device, one launch per device, host adds G totals
        contig      the north-star struct API: a STORM_contiguous_t (filled once with STORM_b200_contig_add_dense) re-homed
                    to G devices (STORM_b200_set_devices + STORM_b200_contig_rehome): first query (rows go up from the
                    pinned mirror in bands, 1/G per PCIe link + NVLink peer pulls) and steady queries
                    (STORM_contig_pairw_intersect_cardinality_blocked)
        wrapper     STORM_wrapper_diag_blocked on a pinned host matrix, H2D + D2H inside the call
grid  the 25 cells of BASELINE config 5 (N 16k .. 256k rows x M 4k .. 1M bits) through STORM_b200_pairw_devices for
      every G: seconds (wall clock of the blocking C call, best of 3), speed-up over G = 1 on the same box, totals equal
      to the single-device total (which tools/bench_configs.py checks against the closed form on every cell).
small the ten shortest cells of that grid with the per-device host threads on and off.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import stormbitmaps_b200 as sb  # noqa: E402

sb.load()
args = sys.argv[1:]
devs = [1, 2, 4, 8]
if "--devices" in args:
    k = args.index("--devices"); devs = [int(x) for x in args[k + 1].split(",")]; del args[k:k + 2]
n_vis = torch.cuda.device_count()
devs = [g for g in devs if g <= n_vis]
SEED = 20260117


def emit(**kw):
    print(json.dumps(kw), flush=True)


def best_wall(fn, reps=3):
    best, val = 1e30, None
    for _ in range(reps):
        t0 = time.perf_counter()
        val = fn()
        best = min(best, time.perf_counter() - t0)
    return best, val


def resident_copies(n, M, gen):
    copies = []
    for d in range(n_vis):
        torch.cuda.set_device(d)
        rows, W = sb.alloc_rows(n, M, device=f"cuda:{d}")
        gen(rows)
        copies.append(rows)
    for d in range(n_vis):
        torch.cuda.synchronize(d)
    torch.cuda.set_device(0)
    return copies, (M + 63) // 64


def run_c3():
    n, M = 200_000, 131_072
    copies, W = resident_copies(n, M, lambda rows: sb.synth_geno_device(rows, M, SEED))
    wp = n * (n - 1) / 2 * W
    ref_total = int(sb.pairw_device(copies[0], n_words=W).item())
    host = torch.empty((n, W), dtype=torch.int64, pin_memory=True)
    host.copy_(copies[0][:, :W])
    torch.cuda.synchronize()
    base = {}
    with sb.StormContiguous(M) as c:
        t0 = time.perf_counter()
        c.add_dense_ptr(host.data_ptr(), n, W)
        emit(config="c3", step="ingest", call="STORM_b200_contig_add_dense", seconds=time.perf_counter() - t0)
        for G in devs:
            sb.pairw_devices(copies[:G], n_words=W)                                   # warm-up (contexts, self-test)
            s, tot = best_wall(lambda: sb.pairw_devices(copies[:G], n_words=W))
            base.setdefault("resident", s)
            emit(config="c3", mode="resident", call="STORM_b200_pairw_devices", devices=G, seconds=s, wp_per_s=wp / s,
                 speedup=base["resident"] / s, match=tot == ref_total)
            sb.set_devices(G)
            c.rehome()
            t0 = time.perf_counter()
            first = c.pairw_intersect_cardinality_blocked(15)
            first_s = time.perf_counter() - t0
            s, tot = best_wall(lambda: c.pairw_intersect_cardinality_blocked(15))
            base.setdefault("contig", s)
            emit(config="c3", mode="contig", call="STORM_contig_pairw_intersect_cardinality_blocked", devices=G, replicas=c.device_count(),
                 first_query_s=first_s, seconds=s, wp_per_s=wp / s, speedup=base["contig"] / s, match=tot == ref_total and first == ref_total)
            sb.wrapper_diag_ptr(host.data_ptr(), n, W, 15)                            # warm-up (scratch arenas)
            s, tot = best_wall(lambda: sb.wrapper_diag_ptr(host.data_ptr(), n, W, 15), reps=2)
            base.setdefault("wrapper", s)
            emit(config="c3", mode="wrapper_e2e", call="STORM_wrapper_diag_blocked (pinned host matrix, H2D + D2H inside)", devices=G,
                 seconds=s, wp_per_s=wp / s, speedup=base["wrapper"] / s, h2d_bytes=n * W * 8, match=tot == ref_total)
        sb.set_device_list(())
    del copies, host
    torch.cuda.empty_cache()


def run_grid():
    for n in (16384, 32768, 65536, 131072, 262144):
        for M in (4096, 16384, 65536, 262144, 1048576):
            copies, W = resident_copies(n, M, lambda rows: sb.synth_geno_device(rows, M, 5))
            wp = n * (n - 1) / 2 * W
            ref_total = int(sb.pairw_device(copies[0], n_words=W).item())
            rec = {"config": "c5", "rows": n, "bits": M, "total": ref_total}
            t1 = None
            for G in devs:
                sb.pairw_devices(copies[:G], n_words=W)
                reps = 3 if wp > 5e12 else 10
                s, tot = best_wall(lambda: sb.pairw_devices(copies[:G], n_words=W), reps=reps)
                t1 = t1 or s
                rec[f"devices{G}"] = {"seconds": s, "wp_per_s": wp / s, "speedup": t1 / s, "efficiency": t1 / s / G, "match": tot == ref_total}
            emit(**rec)
            del copies
            torch.cuda.empty_cache()


def run_small():
    """The short cells of the C5 grid (0.2 .. 8 ms on one device) on every G with the per-device host threads on and
    off (STORM_b200_set_device_threads): where the host's ~12 us of driver calls per device decide the scaling."""
    for n, M in ((16384, 4096), (16384, 16384), (32768, 4096), (16384, 65536), (32768, 16384), (65536, 4096), (16384, 262144),
                 (32768, 65536), (65536, 16384), (131072, 4096)):
        copies, W = resident_copies(n, M, lambda rows: sb.synth_geno_device(rows, M, 5))
        wp = n * (n - 1) / 2 * W
        ref_total = int(sb.pairw_device(copies[0], n_words=W).item())
        rec = {"config": "c5_small", "rows": n, "bits": M, "total": ref_total}
        t1 = None
        for G in devs:
            for threads in ((True,) if G == 1 else (False, True)):
                sb.set_device_threads(threads)
                sb.pairw_devices(copies[:G], n_words=W)
                s, tot = best_wall(lambda: sb.pairw_devices(copies[:G], n_words=W), reps=30)
                t1 = t1 or s
                rec[f"devices{G}" + ("" if threads else "_one_thread")] = {"seconds": s, "speedup": t1 / s, "match": tot == ref_total}
        sb.set_device_threads(True)
        emit(**rec)
        del copies
        torch.cuda.empty_cache()


def run_storm():
    """C2-shaped STORM_t containers (10 000 x 524 288) on device sets: one level the cost model sends to the sparse
    kernels, one it densifies.  A container takes its device set at its first query, so each G gets its own."""
    from oracle import oracle as O
    orc = O.Oracle()
    M, n = 524288, 10000
    for draws in (104, 5242):
        rows = [orc.gen_row_positions(77, i, draws, M) for i in range(n)]
        base, ref_total = None, None
        for G in devs:
            sb.set_devices(G)
            try:
                with sb.Storm() as s:
                    for p in rows:
                        s.add(p)
                    t0 = time.perf_counter()
                    first = s.pairw_intersect_cardinality_blocked(0)
                    first_s = time.perf_counter() - t0
                    sec, tot = best_wall(lambda: s.pairw_intersect_cardinality_blocked(0))
                    ref_total = first if ref_total is None else ref_total
                    base = base or sec
                    emit(config="c2", mode="storm_t", draws=draws, devices=G, route=s.last_route(), first_query_s=first_s, seconds=sec,
                         speedup=base / sec, bitmap_space_wp_per_s=n * (n - 1) / 2 * (M // 64) / sec, match=tot == ref_total and first == ref_total)
            finally:
                sb.set_device_list(())


if __name__ == "__main__":
    emit(devices_visible=n_vis, devices_tested=devs, device=sb.device_info(0))
    for w in (args or ["c3"]):
        {"c3": run_c3, "grid": run_grid, "storm": run_storm, "small": run_small}[w]()
