#!/usr/bin/env python
"""bench.py -- headline benchmark of the all-vs-all intersection-cardinality path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c3|c1|custom --rows R --bits M] [--kernel auto|popc|csa|umma]

Metric (BASELINE.json): 64-bit word-pair AND+popcounts per second,
``wp/s = R(R-1)/2 * ceil(M/64) / seconds`` for the strict upper triangle of XX^T.

A *step* is one full query over the resident matrix.  Default workload is C3
(200,000 rows x 131,072 bits, genotype-like synthetic rows), the configuration
the north-star target is quoted on; it fits one GPU (3.28 GB).

  value   device-resident: rows already in HBM, CUDA events around K steps
          (``STORM_b200_pairw_device``), max over ranks.
  e2e     the same query through the reference-facing host-buffer call
          (``STORM_wrapper_diag_blocked``, storm.h:121-125) from PINNED HOST memory:
          H2D of the whole matrix (in chunks that overlap the kernels) + kernels +
          D2H of the total inside the timed region, every step.  For N>1 every rank
          uploads 1/N of each row band, the bands are all-gathered over NVLink and
          the tiles of the bands that have landed are computed meanwhile
          (stormbitmaps_b200.distributed.pairw_total_from_host).
  N>1     strong scaling: every rank holds the full matrix and owns 1/N of the
          tile raster; the only collective is an 8-byte all-reduce of the total.

``--impl reference`` times the unmodified reference's own CPU kernel
(oracle/_ref/libstorm_ref.so, built from /root/reference by oracle/Makefile) on a
bounded row sample of the same workload, with all host cores (our row-block
partition around the reference's per-pair kernel; the reference itself has no
threading).  Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (rows, bits, generator)
    "c3": (200_000, 131_072, "geno"),
    "c1": (10_000, 65_536, "uniform32768"),
}
SEED = 20260117


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# --------------------------------------------------------------------------- #
# clocks
# --------------------------------------------------------------------------- #
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- #
# reference arm (CPU)
# --------------------------------------------------------------------------- #
def reference_optimal_bsize(W: int) -> int:
    b = int(256e3 / (W * 8))                 # benchmark.cpp:823-824
    return max(5, b)


def run_reference_sample(ref, vals, bsize: int, threads: int) -> float:
    """Upper-triangle total of `vals` with the reference's per-pair kernel, row-block
    partitioned over `threads` host threads (ctypes releases the GIL).  Returns seconds."""
    from concurrent.futures import ThreadPoolExecutor
    n = vals.shape[0]
    TB = max(bsize, (max(1, n // (4 * threads)) // bsize) * bsize) if threads > 1 else n
    tasks = []
    for i0 in range(0, n, TB):
        i1 = min(i0 + TB, n)
        tasks.append(("diag", i0, i1, 0, 0))
        for j0 in range(i1, n, TB):
            tasks.append(("rect", i0, i1, j0, min(j0 + TB, n)))

    def work(t):
        kind, i0, i1, j0, j1 = t
        if kind == "diag":
            return ref.wrapper_diag_blocked(vals[i0:i1], bsize)       # storm.c:222-279
        return ref.rect_blocked(vals, i0, i1, j0, j1, bsize)           # squares of storm.c:1212-1220

    t0 = time.perf_counter()
    if threads == 1:
        total = sum(work(t) for t in tasks)
    else:
        with ThreadPoolExecutor(threads) as ex:
            total = sum(ex.map(work, tasks))
    dt = time.perf_counter() - t0
    run_reference_sample.last_total = total
    return dt


def gen_rows_cpu(orc, gen: str, rows: int, bits: int, threads: int):
    """Sample rows [0, rows) of the workload with the oracle's generator (bit-identical to the device one)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    W = (bits + 63) // 64
    out = np.zeros((rows, W), dtype=np.uint64)
    step = max(1, rows // (threads * 4))

    def fill(r0):
        n = min(step, rows - r0)
        if gen == "geno":
            out[r0:r0 + n] = orc.gen_dense_geno(SEED, n, bits, row0=r0)
        else:
            out[r0:r0 + n] = orc.gen_dense_uniform(SEED, n, int(gen[len("uniform"):]), bits, row0=r0)

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(fill, range(0, rows, step)))
    return out


def cpu_baseline(rows_total: int, bits: int, gen: str, target_s: float, threads: int):
    """Time the reference (or, if its prebuilt .so is absent, the oracle port) on a bounded row sample."""
    from oracle import oracle as O
    orc = O.Oracle()
    W = (bits + 63) // 64
    bsize = reference_optimal_bsize(W)
    if O.have_reference():
        ref, kind = O.Reference(), "reference"
        simd = ref.kernel_name(W)
    else:
        ref, kind, simd = None, "port", "scalar-popcnt"
    # calibrate on a small sample, then size the timed sample for ~target_s
    n0 = min(rows_total, max(4 * bsize, 600))
    vals = gen_rows_cpu(orc, gen, n0, bits, threads)
    if ref is not None:
        dt = run_reference_sample(ref, vals, bsize, threads)
    else:
        t0 = time.perf_counter(); orc.wrapper_diag(vals); dt = time.perf_counter() - t0
    rate = n0 * (n0 - 1) / 2 * W / dt
    n = int(min(rows_total, max(n0, (2 * target_s * rate / W) ** 0.5)))
    n = max(bsize, n // bsize * bsize)
    if n > n0:
        vals = gen_rows_cpu(orc, gen, n, bits, threads)
    return {"orc": orc, "ref": ref, "kind": kind, "simd": simd, "vals": vals, "bsize": bsize, "rows": n, "W": W}


def time_cpu_step(ctx, threads: int) -> float:
    if ctx["ref"] is not None:
        return run_reference_sample(ctx["ref"], ctx["vals"], ctx["bsize"], threads)
    t0 = time.perf_counter()
    ctx["orc"].wrapper_diag(ctx["vals"])
    return time.perf_counter() - t0


def main_reference(args, rows, bits, gen):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0                                           # other ranks exit without work
    threads = os.cpu_count() or 1
    W = (bits + 63) // 64
    per_step = max(2.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    ctx = cpu_baseline(rows, bits, gen, per_step, threads)
    n = ctx["rows"]
    for _ in range(args.warmup):
        time_cpu_step(ctx, threads)
    t = [time_cpu_step(ctx, threads) for _ in range(args.steps)]
    wp = n * (n - 1) / 2 * W
    value = wp * args.steps / sum(t)
    sample = (f"rows [0,{n}) of the {rows}x{bits} workload ({wp:.3e} wp/step), blocked bsize={ctx['bsize']}, "
              f"per-pair kernel={ctx['simd']}, {threads} host threads over row blocks")
    line = {
        "impl": "reference", "metric": "xxt_wordpair_and_popcnt_per_s", "value": value, "unit": "wp/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t) / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: dense {rows}x{bits} XX^T upper triangle", "rows": rows, "bits": bits,
                   "generator": gen, "sample_rows": n},
        "cpu_baseline": {"value": value, "unit": "wp/s", "cores": threads, "kind": ctx["kind"], "sample": sample},
        "e2e": {"value": value, "unit": "wp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- #
# our arm (GPU)
# --------------------------------------------------------------------------- #
def colcount_total_torch(rows_t, W, chunk=16384):
    """sum_k C(c_k,2) from column popcounts with torch ops only (independent checksum)."""
    import torch
    counts = torch.zeros((64, W), dtype=torch.int64, device=rows_t.device)
    for r0 in range(0, rows_t.shape[0], chunk):
        blk = rows_t[r0:r0 + chunk, :W]
        for b in range(64):
            counts[b] += ((blk >> b) & 1).sum(dim=0, dtype=torch.int64)
    return int((counts * (counts - 1) // 2).sum().item())


def sample_tiles(rows: int, tile: int, n: int, seed: int = 7):
    """`n` tiles (bi, bj), bi <= bj, of the tile x tile raster over `rows` rows: the first diagonal tile, its right
    neighbour, the last (ragged) diagonal tile, the last column block against the first and the one before it, the
    rest drawn at random (diagonal and interior)."""
    import random
    nb = (rows + tile - 1) // tile
    picks = [(0, 0), (0, min(1, nb - 1)), (nb - 1, nb - 1), (0, nb - 1), (max(0, nb - 2), nb - 1), (nb // 2, nb // 2)]
    rng = random.Random(seed)
    while len(picks) < n:
        bi = rng.randrange(nb)
        picks.append((bi, rng.randrange(bi, nb)))
    seen, out = set(), []
    for t in picks:
        if t not in seen:
            seen.add(t); out.append(t)
    return out[:n]


def verify_pairs(sb, rows_t, rows: int, bits: int, gen: str, kernel, n_tiles: int, tile: int = 256):
    """Per-pair verification at full size: every pair count of `n_tiles` sampled tile x tile
    rectangles from the resident matrix (STORM_b200_pairw_rect_device, i.e. the timed kernel with per-pair output)
    against the CPU oracle's per-pair kernel on the SAME rows regenerated on the host by the oracle's generator.
    Outside every timed region."""
    import numpy as np
    import torch
    from oracle import oracle as O
    orc = O.Oracle()
    W = (bits + 63) // 64
    t0 = time.perf_counter()
    pairs = bad = 0
    cache = {}

    def host_block(b):
        if b not in cache:
            r0, n = b * tile, min(tile, rows - b * tile)
            cache[b] = (orc.gen_dense_geno(SEED, n, bits, row0=r0) if gen == "geno"
                        else orc.gen_dense_uniform(SEED, n, int(gen[len("uniform"):]), bits, row0=r0))
        return cache[b]

    tiles = sample_tiles(rows, tile, n_tiles)
    for bi, bj in tiles:
        i0, i1, j0, j1 = bi * tile, min(rows, bi * tile + tile), bj * tile, min(rows, bj * tile + tile)
        a, b = host_block(bi), host_block(bj)
        both = np.ascontiguousarray(np.concatenate([a, b]))
        want = orc.rect_counts(both, 0, i1 - i0, i1 - i0, i1 - i0 + (j1 - j0))          # all (i, j) of the rectangle
        if bi == bj:                                                                  # strict upper triangle on the diagonal
            want = np.triu(want, 1)
        got, _ = sb.pairw_rect_device(rows_t, i0, i1, j0, j1, n_words=W, kernel=kernel)
        got = got.cpu().numpy().view(np.uint32)
        bad += int((got != want).sum())
        pairs += (i1 - i0) * (j1 - j0) if bi != bj else (i1 - i0) * (i1 - i0 - 1) // 2
        if len(cache) > 8:
            cache.clear()
    torch.cuda.synchronize()
    return {"tiles_sampled": len(tiles), "tile": tile, "pairs_sampled": int(pairs), "pairs_wrong": int(bad),
            "includes": "first diagonal, last ragged diagonal (rows >= %d), last column block, random diagonal + interior" % ((rows - 1) // tile * tile),
            "checker": "oracle per-pair kernel (oracle/storm_oracle.c) on rows regenerated on the host", "seconds": round(time.perf_counter() - t0, 2)}


def contig_api_bench(rows: int, bits: int, reps: int = 3, bulk: bool = True, timeout: int = 900):
    """The north-star struct API end to end (STORM_contig_new -> rows x STORM_contig_add -> ..._blocked query) through a
    C99 program linked with the library (tools/contig_api_bench.c): ingest, first and steady query seconds."""
    from stormbitmaps_b200 import build as B
    exe = B.tool_path("contig_api_bench")
    if not os.path.isfile(exe):
        return {"unavailable": f"{exe} is not built"}
    try:
        out = subprocess.run([exe, str(bits), str(rows), str(reps), "1" if bulk else "0"], capture_output=True, text=True, timeout=timeout)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else ""
        res = json.loads(line) if line else {"unavailable": out.stderr.strip()[-300:]}
        res["exit_code"] = out.returncode
        return res
    except (subprocess.TimeoutExpired, ValueError, OSError) as e:
        return {"unavailable": repr(e)[:300]}


def ncu_traffic(workload: str, kernel: str, world: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this workload (profiles/traffic.json), or None if none was taken."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(f"{workload}:{kernel}:gpus{world}", {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def main_ours(args, rows, bits, gen):
    import torch
    import torch.distributed as dist
    import stormbitmaps_b200 as sb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sb.load()
    if args.umma_cg:
        sb.set_umma_cta_group(args.umma_cg)
    sb.set_umma_wave_sync(bool(args.wave_sync))
    dev = torch.device("cuda", local_rank)
    W = (bits + 63) // 64
    kernel = args.kernel

    # ---- resident input: every rank generates the same matrix on its own GPU ----
    rows_t, _ = sb.alloc_rows(rows, bits, device=dev)
    if gen == "geno":
        sb.synth_geno_device(rows_t, bits, SEED)
    else:
        sb.synth_uniform_device(rows_t, bits, int(gen[len("uniform"):]), SEED)
    torch.cuda.synchronize()
    wp = rows * (rows - 1) / 2 * W
    total = torch.zeros(1, dtype=torch.int64, device=dev)

    def step():
        total.zero_()
        sb.pairw_device(rows_t, n_words=W, shard=rank, n_shards=world, kernel=kernel, total=total)
        if world > 1:
            dist.all_reduce(total)                         # 8 bytes: the only collective on the path

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sb.set_clock_probe(True)                              # the tensor kernel records its own clock64 / globaltimer deltas
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = sb.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    barrier()
    launches = sb.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    try:
        kclk = sb.last_kernel_clock()                     # of the last timed launch on this rank's device
    except sb.StormError:
        kclk = None
    sb.set_clock_probe(False)
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    elapsed = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    elapsed_ms = float(elapsed.item())
    got_total = int(total.item())

    # ---- e2e: host buffers through the reference-facing call -------------------
    host = torch.empty((rows, W), dtype=torch.int64, pin_memory=True)
    host.copy_(rows_t[:, :W])
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 3))

    from stormbitmaps_b200 import distributed as sbd
    arena = sbd.alloc_stream_arena(rows, W, world, dev, kernel=kernel) if world > 1 else None
    e2e_total_t = torch.zeros(1, dtype=torch.int64, device=dev)

    def e2e_step():
        # N = 1: STORM_wrapper_diag_blocked on the pinned host matrix (chunked H2D overlapped with the kernels, D2H of
        # the total).  N > 1: bands of rows, each uploaded 1/N per rank and all-gathered over NVLink on a side stream
        # while the tiles of the bands already there are being computed; all-reduce, D2H.
        return sbd.pairw_total_from_host(host, kernel=kernel, arena=arena, total=e2e_total_t)

    e2e_total = e2e_step()                                   # warm-up (allocates the scratch arena)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_total = e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = wp * e2e_steps / float(e2e_s.item())

    # the same call on a PAGEABLE buffer (what a C caller's malloc'd matrix is): staged through pinned slots by host threads
    e2e_pageable = None
    if world == 1 and not args.no_extras:
        try:                                                # (an extra: its failure is reported, it does not cost the line)
            import numpy as np
            pageable = np.empty((rows, W), dtype=np.uint64)
            pageable[...] = host.numpy().view(np.uint64)
            sb.wrapper_diag_ptr(pageable.ctypes.data, rows, W, 15)                     # warm-up (staging slots)
            t0 = time.perf_counter()
            got_p = 0
            for _ in range(2):
                got_p = sb.wrapper_diag_ptr(pageable.ctypes.data, rows, W, 15)
            dt = (time.perf_counter() - t0) / 2
            e2e_pageable = {"value": wp / dt, "unit": "wp/s", "seconds": dt, "total": got_p,
                            "call": "STORM_wrapper_diag_blocked on a pageable (numpy / malloc) buffer, H2D + D2H inside"}
            del pageable
        except Exception as e:                              # noqa: BLE001
            e2e_pageable = {"error": repr(e)}

    # in-library multi-device mode (STORM_B200_DEVICES / STORM_b200_set_devices): one process, all visible GPUs behind storm.h
    in_lib = None
    if world == 1 and torch.cuda.device_count() > 1 and not args.no_extras:
        in_lib = {}
        for k in [d for d in (2, 4, 8) if d <= torch.cuda.device_count()]:
            sb.set_devices(k)
            try:
                sb.wrapper_diag_ptr(host.data_ptr(), rows, W, 15)                      # warm-up: arenas on every device
                t0 = time.perf_counter()
                tot_k = sb.wrapper_diag_ptr(host.data_ptr(), rows, W, 15)
                dt = time.perf_counter() - t0
                in_lib[f"devices{k}"] = {"e2e_c_abi": wp / dt, "seconds": dt, "total": tot_k,
                                         "call": "STORM_wrapper_diag_blocked, pinned host buffer, STORM_b200_set_devices(%d)" % k}
            except Exception as e:                          # noqa: BLE001
                in_lib[f"devices{k}"] = {"error": repr(e)}
            finally:
                sb.set_device_list(())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- verification (outside the timed regions) -------------------------------
    closed = colcount_total_torch(rows_t, W)
    ok = (got_total == closed) and (e2e_total == closed)
    if e2e_pageable is not None and "total" in e2e_pageable:
        e2e_pageable["match"] = e2e_pageable["total"] == closed
        ok = ok and e2e_pageable["match"]
    if in_lib:
        for v in in_lib.values():
            if "total" in v:
                v["match"] = v["total"] == closed
                ok = ok and v["match"]
    pairs_check = None
    if args.verify_pairs > 0:
        try:
            pairs_check = verify_pairs(sb, rows_t, rows, bits, gen, kernel, args.verify_pairs)
        except Exception as e:                              # noqa: BLE001  (reported in the line; the run then fails below)
            pairs_check = {"error": repr(e), "pairs_sampled": 0, "pairs_wrong": -1}
        ok = ok and pairs_check["pairs_wrong"] == 0

    peaks, peak_src = load_peaks()
    value = wp * args.steps / (elapsed_ms * 1e-3)
    used_kernel = sb.resolved_kernel_name(kernel, W)
    n_tiles, tm, tn = sb.tile_count(rows, used_kernel)
    # dominant kernel: one launch per step per rank; its duration is the step time
    # minus the 8-byte all-reduce (N>1), measured by the same CUDA events
    launch_ms = elapsed_ms / args.steps
    traffic = ncu_traffic(args.workload, used_kernel, world)
    if used_kernel in ("fp4", "umma"):
        # Algorithmic work: 64 one-bit MACs = 128 ops per 64-bit word pair, carried by tcgen05.mma
        # kind::mxf4 on bits unpacked to E2M1 nibbles (fp32 accumulators holding exact integers) or kind::i8 on bits
        # unpacked to bytes (s32).  MEASURED_PEAKS.json has no fp4 / int8 figure, so the denominator is the tensor
        # pipe itself: 16384 (mxf4) resp. 8192 (i8) MACs per clock and SM -- the rate ncu reports for this kernel at
        # 99 % pipe activity (profiles/r01_fp4_c3_ncu_full_v4.md) and the issue probe below reproduces -- x SMs x the
        # part's MAXIMUM SM clock.  `frac_at_run_clock` divides by the clock the launch actually ran at, measured by
        # the kernel itself (clock64 delta / globaltimer delta per CTA): under the 1 kW cap the delivered clock is
        # below what nvidia-smi samples.
        fp4 = used_kernel == "fp4"
        info = sb.device_info(local_rank)
        sms = info["sm_count"]
        mac_per_clk = 16384.0 if fp4 else 8192.0
        sm_max = float((clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0)
        # is one clock64 tick one SM cycle?  250 ms spin on the otherwise idle device, nvidia-smi sampled beside it
        cal_sampler = ClockSampler(local_rank)
        cal_ticks_mhz = sb.microbench(8)[1]
        cal = cal_sampler.stop()
        smi_idle = cal.get("sm_mhz") or sm_max
        tick_ratio = smi_idle / cal_ticks_mhz if cal_ticks_mhz else 1.0
        if abs(tick_ratio - round(tick_ratio)) < 0.04 and round(tick_ratio) >= 1:
            tick_ratio = float(round(tick_ratio))
        run_mhz = kclk["mhz"] * tick_ratio if kclk else None
        probe_rate, probe_ticks_mhz = sb.microbench(7 if fp4 else 5)
        probe_mac_per_clk = probe_rate / 2.0 / (probe_ticks_mhz * tick_ratio * 1e6) / sms if probe_ticks_mhz else None
        ops_per_launch = wp / world * 128.0
        achieved = ops_per_launch / (launch_ms * 1e-3) / 1e12
        peak = sms * mac_per_clk * 2.0 * sm_max * 1e6 / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "dense_umma_kernel<2, FP4>" if fp4 else "dense_umma_kernel<2, i8>",
                    "unit_note": "tensor ops per second / 1e12 (1 MAC = 2 ops) on %s operands holding bits" % ("E2M1" if fp4 else "u8"),
                    "peak_source": "tensor pipe: %d MAC/clk/SM (%s) x %d SMs x %.0f MHz max SM clock; MEASURED_PEAKS.json has no %s figure "
                                   "(its bf16 burst x %d = %.0f)" % (mac_per_clk, "kind::mxf4" if fp4 else "kind::i8", sms, sm_max,
                                                                     "fp4" if fp4 else "int8", 4 if fp4 else 2,
                                                                     (4 if fp4 else 2) * float(peaks.get("bf16_tflops", 1590.0))),
                    "peak_nominal_dense": 9000.0 if fp4 else 4500.0, "frac_of_nominal": achieved / (9000.0 if fp4 else 4500.0),
                    "issue_probe": {"tops": probe_rate / 1e12, "mac_per_clk_per_sm": probe_mac_per_clk,
                                    "frac_of_pipe": probe_mac_per_clk / mac_per_clk if probe_mac_per_clk else None,
                                    "what": "the kernel's own tcgen05.mma issued back to back on zero operands, cycles counted by the issuing warp (STORM_b200_microbench(%d))" % (7 if fp4 else 5)},
                    "clock64_calibration": {"ticks_per_us_idle_spin": cal_ticks_mhz, "nvidia_smi_sm_mhz_during_spin": smi_idle,
                                            "sm_cycles_per_tick": tick_ratio},
                    "algorithmic": "128 ops per 64-bit word pair x wp per launch",
                    "hbm_gbs_compulsory": rows * W * 8 / (launch_ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"]}
        if run_mhz:
            roofline["kernel_clock_mhz"] = run_mhz
            roofline["kernel_clock_spread_mhz"] = [kclk["min_mhz"] * tick_ratio, kclk["max_mhz"] * tick_ratio]
            roofline["peak_at_run_clock"] = peak * run_mhz / sm_max
            roofline["frac_at_run_clock"] = achieved / (peak * run_mhz / sm_max)
            roofline["mac_per_clk_per_sm"] = ops_per_launch / 2.0 / (launch_ms * 1e-3) / (run_mhz * 1e6) / sms
            roofline["run_clock_source"] = "in-kernel: clock64 delta / globaltimer delta around the persistent loop, mean over CTAs of the last timed launch"
    else:
        # CUDA-core kernels: bound by the POPC issue rate, measured on this device
        popc_rate, _ = sb.microbench(0)
        per_wp = 2.0                                       # 2 POPC.32 per 64-bit word pair (direct form)
        achieved = wp / world * per_wp / (launch_ms * 1e-3)
        roofline = {"bound": "issue", "achieved": achieved / 1e12, "peak": popc_rate / 1e12, "unit": "T POPC.32/s",
                    "frac": achieved / popc_rate, "traffic": traffic, "kernel": "dense_tile_kernel",
                    "peak_source": "STORM_b200_microbench(POPC) on this device (measured)",
                    "algorithmic": "2 LOP3.32 + 2 POPC.32 per 64-bit word pair (SURVEY.md 8(d)); the carry-save kernel "
                                   "issues fewer POPC than that, so its fraction can exceed 1",
                    "hbm_gbs_compulsory": rows * W * 8 / (launch_ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"]}

    dtype = {"fp4": "e2m1 x e2m1 -> f32 accumulators holding exact integers < 2^24 (bits unpacked to FP4 nibbles)",
             "umma": "u8 x u8 -> s32 (bits unpacked to bytes)"}.get(used_kernel, "u64 and + popc")
    line = {
        "metric": "xxt_wordpair_and_popcnt_per_s", "value": value, "unit": "wp/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": f"{args.workload}: dense {rows}x{bits} XX^T upper triangle", "rows": rows, "bits": bits,
                   "generator": gen, "seed": SEED, "kernel": used_kernel, "tile": [tm, tn], "tiles": n_tiles, "wave_sync": bool(args.wave_sync),
                   "parallelism": f"tile-raster shards x{world}, rows replicated",
                   "l2": "inputs larger than L2 (%.2f GB vs 126 MB)" % (rows * W * 8 / 1e9)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "wp/s", "h2d_bytes_per_step": rows * W * 8, "d2h_bytes_per_step": 8,
                "steps": e2e_steps,
                "call": "STORM_wrapper_diag_blocked (host buffer, pinned; upload chunks overlap the kernels)" if world == 1 else
                        "stormbitmaps_b200.distributed.pairw_total_from_host (row bands: 1/N slice H2D per rank + NVLink all-gather, pipelined with the tile kernels; all-reduce)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "verified": {"total": got_total, "closed_form_total": closed, "match": ok,
                     "pairs_sampled": pairs_check["pairs_sampled"] if pairs_check else 0, "pairs": pairs_check},
        "step_ms": step_ms,
    }
    if e2e_pageable is not None:
        line["e2e"]["pageable"] = e2e_pageable
        if "value" in e2e_pageable:
            line["e2e"]["pageable_over_pinned"] = e2e_pageable["value"] / e2e_value
    if in_lib:
        line["in_library_devices"] = in_lib
    if world == 1 and not args.no_extras and args.contig_api_rows != 0:
        n_api = args.contig_api_rows if args.contig_api_rows > 0 else rows
        line["contig_api"] = contig_api_bench(n_api, bits)
    if world == 1 and not args.no_cpu_baseline:
        try:
            threads = 1                                     # the reference is single-threaded
            ctx = cpu_baseline(rows, bits, gen, 12.0, threads)
            dt = time_cpu_step(ctx, threads)
            n = ctx["rows"]
            cwp = n * (n - 1) / 2 * W
            line["cpu_baseline"] = {"value": cwp / dt, "unit": "wp/s", "cores": threads, "kind": ctx["kind"],
                                    "sample": f"rows [0,{n}) of the workload, STORM_wrapper_diag_blocked(bsize={ctx['bsize']}), "
                                              f"per-pair kernel={ctx['simd']}, {dt:.1f} s"}
        except Exception as e:                              # noqa: BLE001
            line["cpu_baseline"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        print(f"VERIFICATION FAILED: total {got_total} e2e {e2e_total} closed form {closed}", file=sys.stderr)
        return 1
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c1", "custom"])
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--bits", type=int, default=None)
    ap.add_argument("--kernel", default="auto", choices=["auto", "popc", "csa", "umma", "fp4", "b1"])
    ap.add_argument("--wave-sync", type=int, default=1, help="UMMA kernel: keep the tiles of a wave in step (L2 reuse); 0 = off")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify-pairs", type=int, default=64, help="tiles of 256 x 256 pairs checked pair by pair against the oracle after the timed region (0 = off)")
    ap.add_argument("--contig-api-rows", type=int, default=-1, help="rows of the struct-API end-to-end measurement (-1 = the workload's, 0 = skip)")
    ap.add_argument("--no-extras", action="store_true", help="skip the pageable / in-library-devices / struct-API measurements")
    ap.add_argument("--umma-cg", type=int, default=0, help="cta_group of the UMMA kernel (1 or 2; 0 = library default)")
    args = ap.parse_args()
    if args.workload == "custom":
        rows, bits, gen = args.rows or 20000, args.bits or 131072, "geno"
    else:
        rows, bits, gen = WORKLOADS[args.workload]
        rows, bits = args.rows or rows, args.bits or bits
    if args.impl == "reference":
        return main_reference(args, rows, bits, gen)
    return main_ours(args, rows, bits, gen)


if __name__ == "__main__":
    sys.exit(main())
