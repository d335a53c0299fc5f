"""Multi-GPU sharding logic on CPU: world_size-2 gloo processes (no GPU needed).

Every rank owns a contiguous range of the tile raster (STORM_b200_shard_tiles);
the only exchange on the path is the all-reduce of the 64-bit partial totals.
Here each rank evaluates its tiles with the CPU oracle (the checker standing in
for the kernel), so the test pins the raster, the shard ranges and the reduce --
the host logic bench.py and a multi-GPU caller rely on.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, kernel, n_rows, M, draws, seed, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import stormbitmaps_b200 as sb
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        orc = O.Oracle()
        vals = orc.gen_dense_uniform(seed, n_rows, draws, M)        # every rank holds the full matrix
        begin, end = sb.shard_tiles(n_rows, rank, world, kernel)
        part, pairs = 0, 0
        for t in range(begin, end):
            i0, i1, j0, j1 = sb.tile_rect(n_rows, t, kernel)
            part += orc.rect_total(vals, i0, i1, j0, j1)             # strict upper triangle of the rectangle
            ii, jj = np.meshgrid(np.arange(i0, i1), np.arange(j0, j1), indexing="ij")
            pairs += int((jj > ii).sum())
        t = torch.tensor([part, pairs, end - begin], dtype=torch.int64)
        dist.all_reduce(t)                                            # the path's only collective
        if rank == 0:
            out.put((int(t[0]), int(t[1]), int(t[2]), orc.wrapper_diag(vals)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kernel,n_rows", [("popc", 700), ("umma", 1100), ("csa", 513)])
def test_two_ranks_partition_the_triangle(kernel, n_rows):
    import torch.multiprocessing as mp
    import stormbitmaps_b200 as sb
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    M, draws, seed, world = 2048, 700, 5, 2
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, kernel, n_rows, M, draws, seed, out)) for r in range(world)]
    for p in procs:
        p.start()
    total, pairs, tiles, exact = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == exact
    assert pairs == n_rows * (n_rows - 1) // 2
    assert tiles == sb.tile_count(n_rows, kernel)[0]


def test_shard_ranges_are_balanced_and_contiguous():
    import stormbitmaps_b200 as sb
    for kernel in ("popc", "umma"):
        for n_rows in (2, 255, 256, 257, 10_000, 200_000):
            n_tiles = sb.tile_count(n_rows, kernel)[0]
            for world in (1, 2, 3, 8):
                prev_end, sizes = 0, []
                for r in range(world):
                    b, e = sb.shard_tiles(n_rows, r, world, kernel)
                    assert b == prev_end and e >= b
                    prev_end = e
                    sizes.append(e - b)
                assert prev_end == n_tiles
                assert max(sizes) - min(sizes) <= 1


def test_tiles_cover_every_pair_once():
    import stormbitmaps_b200 as sb
    for kernel, n_rows in (("popc", 300), ("umma", 600), ("umma", 256), ("popc", 129)):
        seen = np.zeros((n_rows, n_rows), dtype=np.int32)
        for t in range(sb.tile_count(n_rows, kernel)[0]):
            i0, i1, j0, j1 = sb.tile_rect(n_rows, t, kernel)
            seen[i0:i1, j0:j1] += 1
        iu = np.triu_indices(n_rows, k=1)
        assert (seen[iu] == 1).all()


def _gather_main(rank, world, port, n_rows, n_words, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from stormbitmaps_b200 import distributed as D
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1234)                       # every rank holds the same host matrix
        host = torch.randint(-2**62, 2**62, (n_rows, n_words), dtype=torch.int64, generator=g)
        arena = D.alloc_gather_arena(n_rows, n_words, world, "cpu")
        D.gather_rows(host, arena, rank, world)
        ok = bool((arena[:n_rows, :n_words] == host).all()) and bool((arena[n_rows:] == 0).all()) \
            and bool((arena[:, n_words:] == 0).all())
        r0, r1, height = D.slice_bounds(n_rows, rank, world)
        t = torch.tensor([int(ok), r1 - r0], dtype=torch.int64)
        dist.all_reduce(t)
        if rank == 0:
            out.put((int(t[0]), int(t[1]), height, tuple(arena.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,n_words", [(1001, 20), (64, 32), (3, 5)])
def test_two_ranks_gather_row_slices(n_rows, n_words):
    """Host-matrix query at N > 1: each rank uploads its slice, the all-gather replicates the arena."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_gather_main, args=(r, world, port, n_rows, n_words, out)) for r in range(world)]
    for p in procs:
        p.start()
    n_ok, rows_covered, height, shape = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n_ok == world and rows_covered == n_rows
    assert shape == (world * height, (n_words + 15) // 16 * 16)


def test_raster_is_monotone_in_rows():
    """Tiles of raster groups <= g only touch rows below (g + 1) * group_rows (group_rows = the raster's column-block
    group x tile_cols, as STORM_b200_tiles_below_row reports it): what the streamed host-buffer query relies on to
    start computing before the upload has finished."""
    import stormbitmaps_b200 as sb
    for kernel, n_rows in (("umma", 5000), ("popc", 3000), ("umma", 2049)):
        n_tiles, tm, tn = sb.tile_count(n_rows, kernel)
        group_rows = sb.tiles_below_row(n_rows, 0, kernel)[1]
        assert group_rows % tn == 0 and group_rows >= tn
        last_group = 0
        for t in range(n_tiles):
            i0, i1, j0, j1 = sb.tile_rect(n_rows, t, kernel)
            g = j0 // group_rows
            assert g >= last_group, (kernel, t)
            last_group = g
            assert i1 <= min(n_rows, (g + 1) * group_rows) and j1 <= min(n_rows, (g + 1) * group_rows)


# ---- pipelined host query (distributed.pairw_total_from_host, N > 1) ---------------------------------
def test_stream_plan_bands_partition_rows_and_tiles():
    """Bands are contiguous in rows and in raster tiles, every tile of band b reads only rows the bands <= b
    carry, band heights split into `world` equal slices, and the ranks' shares partition a band's tiles."""
    import stormbitmaps_b200 as sb
    from stormbitmaps_b200 import distributed as D
    for kernel, n_rows, world, bands in (("umma", 9000, 2, 8), ("umma", 9000, 3, 4), ("umma", 2049, 8, 8),
                                         ("popc", 5000, 2, 3), ("umma", 300, 2, 8), ("umma", 20000, 8, 5)):
        kid = sb.resolve_kernel(kernel, 1024)
        plan = D.stream_plan(n_rows, world, kid, bands)
        n_tiles = sb.tile_count(n_rows, kernel)[0]
        assert 1 <= len(plan) <= bands
        assert plan[0][0] == 0 and plan[0][2] == 0 and plan[-1][1] == n_rows and plan[-1][3] == n_tiles
        for b, (r0, r1, t0, t1) in enumerate(plan):
            assert r1 > r0 and t1 >= t0
            if b:
                assert r0 == plan[b - 1][1] and t0 == plan[b - 1][3]
            if b < len(plan) - 1:
                assert (r1 - r0) % world == 0
            for t in {t0, (t0 + t1) // 2, max(t0, t1 - 1)} if t1 > t0 else ():
                i0, i1, j0, j1 = sb.tile_rect(n_rows, t, kernel)
                assert i1 <= r1 and j1 <= r1, (kernel, n_rows, b, t)
            if b and t1 > t0:                                   # ... and the band's first tile really needs the band
                i0, i1, j0, j1 = sb.tile_rect(n_rows, t0, kernel)
                assert max(i1, j1) > r0
            prev = t0
            for r in range(world):
                tb, te = D.rank_tiles(t0, t1, r, world)
                assert tb == prev and te >= tb
                prev = te
            assert prev == t1
            # slices of the band: equal heights at fixed places, together exactly the band's rows
            covered = 0
            for r in range(world):
                a, e, h = D.band_slice(r0, r1, r, world)
                assert e - a <= h and (a == r0 + r * h or a == r1)
                covered += e - a
            assert covered == r1 - r0
        assert D.plan_arena_rows(plan, world) >= n_rows


def _band_main(rank, world, port, n_rows, n_words, bands, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import stormbitmaps_b200 as sb
    from stormbitmaps_b200 import distributed as D
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(99)
        host = torch.randint(-2**62, 2**62, (n_rows, n_words), dtype=torch.int64, generator=g)
        kid = sb.resolve_kernel("umma", n_words)
        plan = D.stream_plan(n_rows, world, kid, bands)
        arena = D.alloc_stream_arena(n_rows, n_words, world, "cpu", plan)
        ok = True
        for (r0, r1, t0, t1) in plan:
            D.gather_band(host, arena, r0, r1, rank, world)
            ok &= bool((arena[:r1, :n_words] == host[:r1]).all())     # everything up to this band has landed
            ok &= bool((arena[r1 + world:] == 0).all())               # later bands untouched (bar slice padding)
        ok &= bool((arena[n_rows:] == 0).all()) and bool((arena[:, n_words:] == 0).all())
        t = torch.tensor([int(ok), len(plan)], dtype=torch.int64)
        dist.all_reduce(t)
        if rank == 0:
            out.put((int(t[0]), int(t[1])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,n_words,bands", [(9001, 4, 4), (5000, 16, 8), (100, 3, 8)])
def test_two_ranks_gather_row_bands(n_rows, n_words, bands):
    """Band-wise upload + all-gather of the pipelined host query replicates the matrix band by band."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_band_main, args=(r, world, port, n_rows, n_words, bands, out)) for r in range(world)]
    for p in procs:
        p.start()
    n_ok, n_bands = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n_ok == world and n_bands >= world
