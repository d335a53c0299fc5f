"""Multi-GPU form of the all-vs-all query: one process per GPU, rows replicated, tile-raster shards.

The path shards by independent tiles, so the only data-path collective a
device-resident query needs is the 8-byte all-reduce of the partial totals.  A query that starts from
a HOST matrix has one more exchange worth doing: instead of every rank pulling the whole matrix over
its own PCIe link (N x the bytes, and the PCIe time does not shrink with N), rank r uploads rows
[r N/G, (r+1) N/G) only and the slices are all-gathered over NVLink / NVSwitch (900 GB/s per direction
per GPU against ~55 GB/s of PCIe), after which every rank holds the full arena and runs its shard.

``torch.distributed`` is the plumbing here (process group, NCCL collectives, device memory); the
compute is ``STORM_b200_pairw_device`` of libstorm_b200.so.
"""
from __future__ import annotations

from typing import Optional, Tuple

from . import api


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


def pairw_total_from_host(host_rows, kernel=api.KERNEL_AUTO, group=None, arena=None, total=None) -> int:
    """Upper-triangle intersection total of a host matrix, computed by all ranks of ``group``:
    slice upload + NVLink all-gather + this rank's shard of the tile raster + all-reduce.  Every rank passes
    the same matrix and gets the same total.  With one rank this is ``STORM_wrapper_diag_blocked``."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    n_rows, n_words = host_rows.shape
    if world == 1:
        return api.wrapper_diag_shard_ptr(host_rows.data_ptr(), n_rows, n_words, 0, 1, kernel)
    rank = dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    if arena is None:
        arena = alloc_gather_arena(n_rows, n_words, world, dev)
    if total is None:
        total = torch.zeros(1, dtype=torch.int64, device=dev)
    else:
        total.zero_()
    gather_rows(host_rows, arena, rank, world, group)
    api.pairw_device(arena[:n_rows], n_words=n_words, shard=rank, n_shards=world, kernel=kernel, total=total)
    dist.all_reduce(total, group=group)
    return int(total.item())                                  # D2H of the result
