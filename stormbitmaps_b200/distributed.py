"""Multi-GPU form of the all-vs-all query: one process per GPU, rows replicated, tile-raster shards.

The path shards by independent tiles, so the only data-path collective a
device-resident query needs is the 8-byte all-reduce of the partial totals.  A query that starts from
a HOST matrix has one more exchange worth doing: instead of every rank pulling the whole matrix over
its own PCIe link (N x the bytes, and the PCIe time does not shrink with N), rank r uploads rows
[r N/G, (r+1) N/G) only and the slices are all-gathered over NVLink / NVSwitch (900 GB/s per direction
per GPU against ~55 GB/s of PCIe), after which every rank holds the full arena and runs its shard.

The upload, the all-gather and the kernels are pipelined (``pairw_total_from_host``, N > 1): the rows are
cut into a few bands, band b is uploaded (1/G of it per rank) and all-gathered on a side stream, and as soon
as it has landed every rank launches its 1/G of the raster tiles that only read rows of bands <= b -- the
raster is monotone in the largest row a tile reads (``STORM_b200_tiles_below_row``).  The persistent tile
kernel leaves a couple of SMs free while transfers are still in flight (``STORM_b200_set_umma_reserved_sms``)
so that the all-gather's CTAs find a place to run beside it.

``torch.distributed`` is the plumbing here (process group, NCCL collectives, device memory); the
compute is ``STORM_b200_pairw_device`` / ``STORM_b200_pairw_tiles_device`` of libstorm_b200.so.
"""
from __future__ import annotations

from math import gcd
from typing import List, Optional, Tuple

from . import api

# Knobs of the pipelined host query (bench.py / tools A/B them): bands the rows are cut into, and the SMs the
# tile kernel leaves to the all-gather while later bands are still in flight.
STREAM_BANDS = 8
STREAM_RESERVED_SMS = 2


def slice_bounds(n_rows: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Rows [r0, r1) that ``rank`` uploads, and the per-rank slice height (the arena holds
    ``world * height`` rows; rows past ``n_rows`` stay zero and pair with nothing)."""
    height = (n_rows + world - 1) // world
    r0 = min(n_rows, rank * height)
    r1 = min(n_rows, r0 + height)
    return r0, r1, height


def alloc_gather_arena(n_rows: int, n_words: int, world: int, device):
    """Zeroed arena of ``world * height`` rows in the library's layout (row stride padded to 128 bytes)."""
    import torch
    _, _, height = slice_bounds(n_rows, 0, world)
    stride = (n_words + 15) // 16 * 16
    return torch.zeros((world * height, stride), dtype=torch.int64, device=device)


def gather_rows(host_rows, arena, rank: int, world: int, group=None):
    """Upload this rank's slice of ``host_rows`` (pinned host tensor, [n_rows, n_words] int64) into its place in
    ``arena`` and all-gather the slices, so that every rank ends up with every row.  Asynchronous on the
    current stream for CUDA arenas."""
    import torch
    import torch.distributed as dist
    n_rows, n_words = host_rows.shape
    r0, r1, height = slice_bounds(n_rows, rank, world)
    mine = arena[rank * height:(rank + 1) * height]
    if r1 > r0:
        mine[: r1 - r0, :n_words].copy_(host_rows[r0:r1], non_blocking=True)
    if world > 1:
        dist.all_gather_into_tensor(arena.view(-1), mine.reshape(-1), group=group)
    return arena


def stream_plan(n_rows: int, world: int, kernel, max_bands: int = STREAM_BANDS) -> List[Tuple[int, int, int, int]]:
    """Bands of the pipelined host query: ``[(r0, r1, t0, t1), ...]`` -- rows [r0, r1) travel in band b, and
    raster tiles [t0, t1) are the ones that read rows of bands <= b only (so they may run once band b has
    landed).  Band heights are multiples of the raster's row granularity and of ``world`` (except the last
    band), so that a band splits into ``world`` equal slices at fixed places of the arena.  Host-only."""
    n_tiles, band_rows = api.tiles_below_row(n_rows, n_rows, kernel)
    unit = band_rows * world // gcd(band_rows, world)
    n_units = (n_rows + unit - 1) // unit
    n_bands = max(1, min(max_bands, n_units))
    plan, t_prev = [], 0
    for b in range(n_bands):
        r0 = (n_units * b // n_bands) * unit
        r1 = min(n_rows, (n_units * (b + 1) // n_bands) * unit)
        t1 = api.tiles_below_row(n_rows, r1, kernel)[0]
        plan.append((r0, r1, t_prev, t1))
        t_prev = t1
    assert t_prev == n_tiles and plan[-1][1] == n_rows
    return plan


def band_slice(r0: int, r1: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Rows [a, b) of band [r0, r1) that ``rank`` uploads and the per-rank slice height of the band."""
    height = (r1 - r0 + world - 1) // world
    a = min(r1, r0 + rank * height)
    return a, min(r1, a + height), height


def plan_arena_rows(plan, world: int) -> int:
    r0, r1, _, _ = plan[-1]
    return r0 + world * band_slice(r0, r1, 0, world)[2]


def alloc_stream_arena(n_rows: int, n_words: int, world: int, device, plan=None, kernel=api.KERNEL_AUTO):
    """Zeroed arena for the pipelined query: the rows plus the padding of the last band's slices."""
    import torch
    plan = plan or stream_plan(n_rows, world, api.resolve_kernel(kernel, n_words))
    stride = (n_words + 15) // 16 * 16
    return torch.zeros((plan_arena_rows(plan, world), stride), dtype=torch.int64, device=device)


def gather_band(host_rows, arena, r0: int, r1: int, rank: int, world: int, group=None):
    """Upload this rank's slice of band [r0, r1) into its place in ``arena`` and all-gather the band.
    Asynchronous on the current stream for CUDA arenas."""
    import torch.distributed as dist
    n_words = host_rows.shape[1]
    a, b, height = band_slice(r0, r1, rank, world)
    region = arena[r0:r0 + world * height]
    mine = region[rank * height:(rank + 1) * height]
    if b > a:
        mine[: b - a, :n_words].copy_(host_rows[a:b], non_blocking=True)
    if world > 1:
        dist.all_gather_into_tensor(region.view(-1), mine.reshape(-1), group=group)


def rank_tiles(t0: int, t1: int, rank: int, world: int) -> Tuple[int, int]:
    """This rank's contiguous share of the tile range [t0, t1): sizes differ by at most one tile."""
    q, r = divmod(t1 - t0, world)
    begin = t0 + rank * q + min(rank, r)
    return begin, begin + q + (1 if rank < r else 0)


class _StreamState:
    """Side stream and per-band events of the pipelined query, kept per device between calls."""
    by_device = {}

    def __init__(self, n_events: int):
        import torch
        self.comm = torch.cuda.Stream()
        self.events = [torch.cuda.Event() for _ in range(n_events)]

    @classmethod
    def get(cls, device, n_events: int):
        st = cls.by_device.get(device)
        if st is None or len(st.events) < n_events:
            st = cls.by_device[device] = cls(n_events)
        return st


def pairw_total_from_host(host_rows, kernel=api.KERNEL_AUTO, group=None, arena=None, total=None,
                          pipelined: bool = True, reserved_sms: Optional[int] = None,
                          bands: Optional[int] = None) -> int:
    """Upper-triangle intersection total of a host matrix, computed by all ranks of ``group``.  Every rank
    passes the same (pinned) matrix and gets the same total.  With one rank this is
    ``STORM_wrapper_diag_blocked``.  N > 1, ``pipelined`` (default): bands of rows are uploaded (1/N per rank)
    and all-gathered over NVLink on a side stream while the tiles of the bands that have landed are already
    being computed; else: slice upload, one all-gather, this rank's shard, in sequence.  Ends with the 8-byte
    all-reduce and the D2H read of the total."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    n_rows, n_words = host_rows.shape
    if world == 1:
        return api.wrapper_diag_shard_ptr(host_rows.data_ptr(), n_rows, n_words, 0, 1, kernel)
    rank = dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    if total is None:
        total = torch.zeros(1, dtype=torch.int64, device=dev)
    else:
        total.zero_()
    if not pipelined:
        if arena is None:
            arena = alloc_gather_arena(n_rows, n_words, world, dev)
        gather_rows(host_rows, arena, rank, world, group)
        api.pairw_device(arena[:n_rows], n_words=n_words, shard=rank, n_shards=world, kernel=kernel, total=total)
    else:
        kid = api.resolve_kernel(kernel, n_words)
        plan = stream_plan(n_rows, world, kid, bands or STREAM_BANDS)
        if arena is None or arena.shape[0] < plan_arena_rows(plan, world):
            arena = alloc_stream_arena(n_rows, n_words, world, dev, plan)
        st = _StreamState.get(dev, len(plan))
        cur = torch.cuda.current_stream()
        st.comm.wait_stream(cur)                  # an earlier query's kernels may still be reading the arena
        reserve = STREAM_RESERVED_SMS if reserved_sms is None else reserved_sms
        rows_view = arena[:n_rows]
        for b, (r0, r1, t0, t1) in enumerate(plan):
            with torch.cuda.stream(st.comm):
                gather_band(host_rows, arena, r0, r1, rank, world, group)
                st.events[b].record(st.comm)
            cur.wait_event(st.events[b])
            tb, te = rank_tiles(t0, t1, rank, world)
            # per launch, not a process-wide knob: the last (largest) launch has nothing left in flight beside it and takes every SM
            api.pairw_tiles_device(rows_view, tb, te, n_words=n_words, kernel=kid, total=total,
                                   reserved_sms=0 if b == len(plan) - 1 else reserve)
    dist.all_reduce(total, group=group)
    return int(total.item())                                  # D2H of the result
