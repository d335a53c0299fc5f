"""Host-side mirror of the reference's C API for the pairwise path.

The method names follow storm.h one to one (``STORM_contig_add`` ->
``StormContiguous.add`` ...), with the same argument meaning and error
behaviour, so the parity tests read like calls against the reference.  All work
happens in libstorm_b200.so; this module only marshals numpy / torch buffers to
plain pointers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import KERNEL_AUTO, KERNEL_NAMES, StormError, UINT64_MAX, u32p, u64p


def _kernel_id(kernel) -> int:
    return KERNEL_NAMES[kernel] if isinstance(kernel, str) else int(kernel)


def _u32(values) -> np.ndarray:
    return np.ascontiguousarray(values, dtype=np.uint32)


def _query(value: int, what: str) -> int:
    """Reference convention: (uint64)-1 / -2 / -3 are error sentinels (storm.c:1150,1245-1246)."""
    if value >= UINT64_MAX - 2:
        raise StormError(f"{what} returned the error sentinel {value - 2**64}: {_lib.last_error()}")
    return int(value)


class StormContiguous:
    """``STORM_contiguous_t`` (storm.h:188-200): dense rows of ``vector_length`` bits."""

    def __init__(self, vector_length: int):
        self._L = _lib.load()
        self.vector_length = int(vector_length)
        self._h = self._L.STORM_contig_new(self.vector_length)            # STORM_contig_new
        if not self._h:
            raise MemoryError("STORM_contig_new returned NULL")

    # -- construction ------------------------------------------------------
    def add(self, values: Sequence[int]) -> int:
        """``STORM_contig_add``: one row from a sorted position list; returns n_values (0 adds no row)."""
        v = _u32(values)
        rc = self._L.STORM_contig_add(self._h, v.ctypes.data_as(u32p), v.size)
        if rc < 0:
            raise StormError(f"STORM_contig_add failed ({rc}): {_lib.last_error()}")
        return rc

    def add_bulk(self, positions: np.ndarray, offsets: np.ndarray) -> None:
        """``STORM_b200_contig_add_bulk``: many rows at once, bits scattered on the device."""
        p = _u32(positions)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        _lib.check(self._L.STORM_b200_contig_add_bulk(self._h, p.ctypes.data_as(u32p), o.ctypes.data_as(u64p),
                                                      o.size - 1), "STORM_b200_contig_add_bulk")

    def clear(self) -> int:
        return self._L.STORM_contig_clear(self._h)                          # STORM_contig_clear

    # -- queries -----------------------------------------------------------
    def pairw_intersect_cardinality(self) -> int:
        return _query(self._L.STORM_contig_pairw_intersect_cardinality(self._h), "STORM_contig_pairw_intersect_cardinality")

    def pairw_intersect_cardinality_blocked(self, bsize: int = 0) -> int:
        return _query(self._L.STORM_contig_pairw_intersect_cardinality_blocked(self._h, bsize),
                      "STORM_contig_pairw_intersect_cardinality_blocked")

    def pairw_intersect_cardinality_list(self) -> int:
        return _query(self._L.STORM_contig_pairw_intersect_cardinality_list(self._h), "STORM_contig_pairw_intersect_cardinality_list")

    def pairw_intersect_cardinality_blocked_list(self, bsize: int = 0) -> int:
        return _query(self._L.STORM_contig_pairw_intersect_cardinality_blocked_list(self._h, bsize),
                      "STORM_contig_pairw_intersect_cardinality_blocked_list")

    def pairw_shard(self, shard: int, n_shards: int, kernel=KERNEL_AUTO) -> int:
        """Partial total of one shard of the tile raster (multi-GPU: one shard per rank)."""
        return _query(self._L.STORM_b200_contig_pairw_shard(self._h, shard, n_shards, _kernel_id(kernel)),
                      "STORM_b200_contig_pairw_shard")

    def pairw_rect(self, i0: int, i1: int, j0: int, j1: int) -> np.ndarray:
        """Per-pair counts of rows [i0,i1) x [j0,j1), strict upper triangle (j <= i reads 0)."""
        out = np.zeros((i1 - i0, j1 - j0), dtype=np.uint32)
        _lib.check(self._L.STORM_b200_contig_pairw_rect(self._h, i0, i1, j0, j1, out.ctypes.data_as(u32p)),
                   "STORM_b200_contig_pairw_rect")
        return out

    def add_dense(self, rows: np.ndarray) -> None:
        """``STORM_b200_contig_add_dense``: rows given as bitmaps (2-D uint64, one row per line)."""
        v = np.ascontiguousarray(rows, dtype=np.uint64)
        _lib.check(self._L.STORM_b200_contig_add_dense(self._h, v.ctypes.data_as(u64p), v.shape[0], v.shape[1]),
                   "STORM_b200_contig_add_dense")

    def add_dense_ptr(self, ptr: int, n_rows: int, pitch_words: int) -> None:
        _lib.check(self._L.STORM_b200_contig_add_dense(self._h, C.cast(ptr, u64p), n_rows, pitch_words), "STORM_b200_contig_add_dense")

    def rehome(self) -> None:
        """``STORM_b200_contig_rehome``: rebuild the device replicas on the device set in force now."""
        _lib.check(self._L.STORM_b200_contig_rehome(self._h), "STORM_b200_contig_rehome")

    def device_count(self) -> int:
        """Device replicas this container's queries run on (0 before its first use of a device)."""
        return int(self._L.STORM_b200_contig_device_count(self._h))

    def last_list_route(self) -> str:
        """Which kernels answered the last *_list query: 'tile', 'probe' (tile + probe) or 'stream'."""
        return {0: "none", 1: "tile", 2: "probe", 3: "stream"}[self._L.STORM_b200_contig_last_list_route(self._h)]

    def invalidate_device(self) -> None:
        _lib.check(self._L.STORM_b200_contig_invalidate_device(self._h), "STORM_b200_contig_invalidate_device")

    def last_timing(self) -> dict:
        t = (C.c_double * 3)()
        _lib.check(self._L.STORM_b200_contig_last_timing(self._h, t), "STORM_b200_contig_last_timing")
        return {"upload_s": t[0], "kernel_s": t[1], "total_s": t[2]}

    # -- lifetime ----------------------------------------------------------
    def free(self) -> None:
        if self._h:
            self._L.STORM_contig_free(self._h)                              # STORM_contig_free
            self._h = None

    close = free

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Storm:
    """``STORM_t`` (storm.h:175-178): rows of 65536-bit blocks, bitmap or u16 list each."""

    def __init__(self):
        self._L = _lib.load()
        self._h = self._L.STORM_new()                                      # STORM_new
        if not self._h:
            raise MemoryError("STORM_new returned NULL")

    def add(self, values: Sequence[int]) -> int:
        v = _u32(values)
        rc = self._L.STORM_add(self._h, v.ctypes.data_as(u32p), v.size)    # STORM_add
        if rc < 0:
            raise StormError(f"STORM_add failed ({rc}): {_lib.last_error()}")
        return rc

    def clear(self) -> int:
        return self._L.STORM_clear(self._h)

    def pairw_intersect_cardinality(self) -> int:
        return _query(self._L.STORM_pairw_intersect_cardinality(self._h), "STORM_pairw_intersect_cardinality")

    def pairw_intersect_cardinality_blocked(self, bsize: int = 0) -> int:
        return _query(self._L.STORM_pairw_intersect_cardinality_blocked(self._h, bsize),
                      "STORM_pairw_intersect_cardinality_blocked")

    def last_route(self) -> str:
        """Which kernel family answered the last whole-container query: 'sparse', 'dense' or 'split'."""
        return {0: "none", 1: "sparse", 2: "dense", 3: "dense", 4: "split"}[self._L.STORM_b200_storm_last_route(self._h)]

    def pairw_shard(self, shard: int, n_shards: int) -> int:
        return _query(self._L.STORM_b200_storm_pairw_shard(self._h, shard, n_shards), "STORM_b200_storm_pairw_shard")

    def pairw_rect(self, i0: int, i1: int, j0: int, j1: int) -> np.ndarray:
        out = np.zeros((i1 - i0, j1 - j0), dtype=np.uint32)
        _lib.check(self._L.STORM_b200_storm_pairw_rect(self._h, i0, i1, j0, j1, out.ctypes.data_as(u32p)),
                   "STORM_b200_storm_pairw_rect")
        return out

    def intersect_cardinality_square(self, other: "Storm") -> int:
        return _query(self._L.STORM_intersect_cardinality_square(self._h, other._h), "STORM_intersect_cardinality_square")

    def serialized_size(self) -> int:
        return int(self._L.STORM_serialized_size(self._h))

    def free(self) -> None:
        if self._h:
            self._L.STORM_free(self._h)
            self._h = None

    close = free

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# raw host buffers (storm.h:95-148)
# ---------------------------------------------------------------------------
OPS = {"intersect": 0, "union": 1, "diff": 2}


def _compute_func(op):
    """The STORM_compute_func a C caller would pass for a set operation (libalgebra.h:3094-3236)."""
    L = _lib.load()
    getter = {"intersect": L.STORM_get_intersect_count_func, "union": L.STORM_get_union_count_func,
              "diff": L.STORM_get_diff_count_func}[op]
    return getter(0)


def wrapper_diag(vals: np.ndarray, op: str = "intersect") -> int:
    """``STORM_wrapper_diag``: caller-owned host matrix (n_vectors, n_ints) of uint64 -> total.
    ``op`` picks the per-pair kernel pointer handed to it, exactly as a C caller would
    (``STORM_get_{intersect,union,diff}_count_func``)."""
    L = _lib.load()
    v = np.ascontiguousarray(vals, dtype=np.uint64)
    f = None if op == "intersect" else _compute_func(op)
    return _query(L.STORM_wrapper_diag(v.shape[0], v.ctypes.data_as(u64p), v.shape[1], f), "STORM_wrapper_diag")


def wrapper_diag_list(vals: np.ndarray, rows_positions, cutoff: int, bsize: Optional[int] = None) -> int:
    """``STORM_wrapper_diag_list[_blocked]`` (storm.h:104-119, 127-148): host matrix + the caller-built position arrays
    the reference uses to pick a cheaper per-pair code path (n_alts, concatenated positions, offsets)."""
    L = _lib.load()
    v = np.ascontiguousarray(vals, dtype=np.uint64)
    n_alts = np.asarray([len(p) for p in rows_positions], dtype=np.uint32)
    offs = np.zeros(len(rows_positions), dtype=np.uint32)
    if len(rows_positions) > 1:
        offs[1:] = np.cumsum(n_alts[:-1], dtype=np.uint64).astype(np.uint32)
    flat = (np.concatenate([np.asarray(p, dtype=np.uint32) for p in rows_positions]) if len(rows_positions) else np.zeros(0, np.uint32))
    flat = np.ascontiguousarray(flat, dtype=np.uint32)
    f = _compute_func("intersect")
    if bsize is None:
        return _query(L.STORM_wrapper_diag_list(v.shape[0], v.ctypes.data_as(u64p), v.shape[1], n_alts.ctypes.data_as(u32p),
                                                flat.ctypes.data_as(u32p), offs.ctypes.data_as(u32p), f, None, cutoff), "STORM_wrapper_diag_list")
    return _query(L.STORM_wrapper_diag_list_blocked(v.shape[0], v.ctypes.data_as(u64p), v.shape[1], n_alts.ctypes.data_as(u32p),
                                                    flat.ctypes.data_as(u32p), offs.ctypes.data_as(u32p), f, None, cutoff, bsize),
                  "STORM_wrapper_diag_list_blocked")


def wrapper_diag_ptr(ptr: int, n_vectors: int, n_ints: int, bsize: int = 0) -> int:
    """``STORM_wrapper_diag_blocked`` on a raw host pointer (e.g. a pinned torch tensor)."""
    L = _lib.load()
    return _query(L.STORM_wrapper_diag_blocked(n_vectors, C.cast(ptr, u64p), n_ints, None, bsize), "STORM_wrapper_diag_blocked")


def wrapper_diag_shard_ptr(ptr: int, n_vectors: int, n_ints: int, shard: int = 0, n_shards: int = 1,
                           kernel=KERNEL_AUTO) -> int:
    """``STORM_b200_wrapper_diag_shard``: host buffer -> upload -> partial total of one shard."""
    L = _lib.load()
    return _query(L.STORM_b200_wrapper_diag_shard(n_vectors, C.cast(ptr, u64p), n_ints, shard, n_shards,
                                                  _kernel_id(kernel)), "STORM_b200_wrapper_diag_shard")


def resolved_kernel_name(kernel, n_words: int) -> str:
    """Name of the kernel AUTO / the process default resolves to for rows of n_words."""
    kid = _lib.load().STORM_b200_resolve_kernel(_kernel_id(kernel), n_words)
    return {v: k for k, v in KERNEL_NAMES.items()}[kid]


def wrapper_square(v1: np.ndarray, v2: np.ndarray, op: str = "intersect") -> int:
    L = _lib.load()
    a = np.ascontiguousarray(v1, dtype=np.uint64)
    b = np.ascontiguousarray(v2, dtype=np.uint64)
    if a.shape[1] != b.shape[1]:
        raise ValueError("row widths differ")
    return _query(L.STORM_wrapper_square(a.shape[0], a.ctypes.data_as(u64p), b.shape[0], b.ctypes.data_as(u64p),
                                         a.shape[1], None if op == "intersect" else _compute_func(op)), "STORM_wrapper_square")


# ---------------------------------------------------------------------------
# device-resident rows (torch tensors are only pointer carriers here)
# ---------------------------------------------------------------------------
def _rows_args(rows):
    """rows: 2-D torch int64/uint64 CUDA tensor, row-major.  Returns (ptr, n_rows, stride_words)."""
    if rows.dim() != 2 or rows.element_size() != 8 or not rows.is_cuda or rows.stride(1) != 1:
        raise ValueError("rows must be a 2-D CUDA tensor of 64-bit words, contiguous along dim 1")
    return rows.data_ptr(), rows.shape[0], rows.stride(0)


def _stream_handle(stream) -> Optional[int]:
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    return int(stream)


def pairw_device(rows, n_words: Optional[int] = None, shard: int = 0, n_shards: int = 1, kernel=KERNEL_AUTO,
                 total=None, stream=None):
    """``STORM_b200_pairw_device``: accumulate this shard's partial total into ``total`` (1-elem int64 CUDA tensor).

    Asynchronous on ``stream`` (default: torch's current stream).  Returns ``total``."""
    import torch
    L = _lib.load()
    ptr, n_rows, stride = _rows_args(rows)
    if total is None:
        total = torch.zeros(1, dtype=torch.int64, device=rows.device)
    _lib.check(L.STORM_b200_pairw_device(ptr, n_rows, n_words or rows.shape[1], stride, shard, n_shards,
                                         _kernel_id(kernel), total.data_ptr(), _stream_handle(stream)),
               "STORM_b200_pairw_device")
    return total


def pairw_tiles_device(rows, tile_begin: int, tile_end: int, n_words: Optional[int] = None, kernel=KERNEL_AUTO,
                       total=None, stream=None, reserved_sms: Optional[int] = None):
    """``STORM_b200_pairw_tiles_device[_ex]``: accumulate the partial total of raster tiles [tile_begin, tile_end).
    ``reserved_sms`` (per launch): SMs the persistent tensor kernel leaves to a collective running beside it."""
    import torch
    L = _lib.load()
    ptr, n_rows, stride = _rows_args(rows)
    if total is None:
        total = torch.zeros(1, dtype=torch.int64, device=rows.device)
    if reserved_sms is None:
        _lib.check(L.STORM_b200_pairw_tiles_device(ptr, n_rows, n_words or rows.shape[1], stride, tile_begin, tile_end,
                                                   _kernel_id(kernel), total.data_ptr(), _stream_handle(stream)),
                   "STORM_b200_pairw_tiles_device")
    else:
        _lib.check(L.STORM_b200_pairw_tiles_device_ex(ptr, n_rows, n_words or rows.shape[1], stride, tile_begin, tile_end,
                                                      _kernel_id(kernel), int(reserved_sms), total.data_ptr(), _stream_handle(stream)),
                   "STORM_b200_pairw_tiles_device_ex")
    return total


def pairw_rect_device(rows, i0, i1, j0, j1, n_words: Optional[int] = None, strict_upper: bool = True,
                      kernel=KERNEL_AUTO, want_counts: bool = True, stream=None):
    """``STORM_b200_pairw_rect_device``: (counts[int32 view of uint32], total) of a rectangle of pairs."""
    import torch
    L = _lib.load()
    ptr, n_rows, stride = _rows_args(rows)
    out = torch.zeros((i1 - i0, j1 - j0), dtype=torch.int32, device=rows.device) if want_counts else None
    total = torch.zeros(1, dtype=torch.int64, device=rows.device)
    _lib.check(L.STORM_b200_pairw_rect_device(ptr, n_rows, n_words or rows.shape[1], stride, i0, i1, j0, j1,
                                              int(strict_upper), _kernel_id(kernel),
                                              out.data_ptr() if want_counts else None, j1 - j0,
                                              total.data_ptr(), _stream_handle(stream)),
               "STORM_b200_pairw_rect_device")
    return out, total


def pairw_op_device(rows, op: str, n_words: Optional[int] = None, kernel=KERNEL_AUTO, stream=None):
    """``STORM_b200_pairw_op_device``: upper-triangle total of |a&b|, |a|b| or |a^b| (1-elem int64 CUDA tensor)."""
    import torch
    L = _lib.load()
    ptr, n_rows, stride = _rows_args(rows)
    total = torch.zeros(1, dtype=torch.int64, device=rows.device)
    _lib.check(L.STORM_b200_pairw_op_device(ptr, n_rows, n_words or rows.shape[1], stride, OPS[op], _kernel_id(kernel),
                                            total.data_ptr(), _stream_handle(stream)), "STORM_b200_pairw_op_device")
    return total


def pairw_rect_op_device(rows, op: str, i0, i1, j0, j1, n_words: Optional[int] = None, strict_upper: bool = True,
                         kernel=KERNEL_AUTO, stream=None):
    """``STORM_b200_pairw_rect_op_device``: (counts, total) of a rectangle of pairs under a set operation."""
    import torch
    L = _lib.load()
    ptr, n_rows, stride = _rows_args(rows)
    out = torch.zeros((i1 - i0, j1 - j0), dtype=torch.int32, device=rows.device)
    total = torch.zeros(1, dtype=torch.int64, device=rows.device)
    _lib.check(L.STORM_b200_pairw_rect_op_device(ptr, n_rows, n_words or rows.shape[1], stride, i0, i1, j0, j1,
                                                 int(strict_upper), OPS[op], _kernel_id(kernel), out.data_ptr(), j1 - j0,
                                                 total.data_ptr(), _stream_handle(stream)), "STORM_b200_pairw_rect_op_device")
    return out, total


def row_popcounts_device(rows, n_words: Optional[int] = None, stream=None):
    """``STORM_b200_row_popcounts_device``: set bits per row (int32 CUDA tensor)."""
    import torch
    L = _lib.load()
    ptr, n_rows, stride = _rows_args(rows)
    out = torch.zeros(n_rows, dtype=torch.int32, device=rows.device)
    _lib.check(L.STORM_b200_row_popcounts_device(ptr, n_rows, n_words or rows.shape[1], stride, out.data_ptr(),
                                                 _stream_handle(stream)), "STORM_b200_row_popcounts_device")
    return out


def square_device(rows1, rows2, n_words: Optional[int] = None, kernel=KERNEL_AUTO, want_counts: bool = False, stream=None):
    """``STORM_b200_square_device``: XY^T over two device matrices."""
    import torch
    L = _lib.load()
    p1, n1, s1 = _rows_args(rows1)
    p2, n2, s2 = _rows_args(rows2)
    out = torch.zeros((n1, n2), dtype=torch.int32, device=rows1.device) if want_counts else None
    total = torch.zeros(1, dtype=torch.int64, device=rows1.device)
    _lib.check(L.STORM_b200_square_device(p1, n1, s1, p2, n2, s2, n_words or rows1.shape[1], _kernel_id(kernel),
                                          out.data_ptr() if want_counts else None, n2, total.data_ptr(),
                                          _stream_handle(stream)), "STORM_b200_square_device")
    return out, total


def alloc_rows(n_rows: int, M: int, device="cuda"):
    """Zeroed device arena in the library's layout: row stride padded to 128 bytes."""
    import torch
    n_words = (M + 63) // 64
    stride = (n_words + 15) // 16 * 16
    return torch.zeros((n_rows, stride), dtype=torch.int64, device=device), n_words


def synth_uniform_device(rows, M: int, n_draws: int, seed: int, row0: int = 0, stream=None):
    L = _lib.load()
    ptr, n_rows, stride = _rows_args(rows)
    _lib.check(L.STORM_b200_synth_uniform_device(ptr, n_rows, (M + 63) // 64, stride, M, n_draws, seed, row0,
                                                 _stream_handle(stream)), "STORM_b200_synth_uniform_device")
    return rows


def synth_geno_device(rows, M: int, seed: int, row0: int = 0, stream=None):
    L = _lib.load()
    ptr, n_rows, stride = _rows_args(rows)
    _lib.check(L.STORM_b200_synth_geno_device(ptr, n_rows, (M + 63) // 64, stride, M, seed, row0,
                                              _stream_handle(stream)), "STORM_b200_synth_geno_device")
    return rows


def tile_count(n_rows: int, kernel=KERNEL_AUTO):
    L = _lib.load()
    tm, tn = C.c_uint32(), C.c_uint32()
    n = L.STORM_b200_tile_count(n_rows, _kernel_id(kernel), C.byref(tm), C.byref(tn))
    return int(n), tm.value, tn.value


def shard_tiles(n_rows: int, shard: int, n_shards: int, kernel=KERNEL_AUTO):
    """``STORM_b200_shard_tiles`` (host only): tile range [begin, end) owned by one shard."""
    L = _lib.load()
    b, e = C.c_uint64(), C.c_uint64()
    _lib.check(L.STORM_b200_shard_tiles(n_rows, _kernel_id(kernel), shard, n_shards, C.byref(b), C.byref(e)),
               "STORM_b200_shard_tiles")
    return int(b.value), int(e.value)


def tiles_below_row(n_rows: int, row_limit: int, kernel=KERNEL_AUTO):
    """``STORM_b200_tiles_below_row`` (host only): (number of leading raster tiles that read only rows below
    ``row_limit``, row granularity at which that number grows)."""
    L = _lib.load()
    t, band = C.c_uint64(), C.c_uint64()
    _lib.check(L.STORM_b200_tiles_below_row(n_rows, _kernel_id(kernel), row_limit, C.byref(t), C.byref(band)),
               "STORM_b200_tiles_below_row")
    return int(t.value), int(band.value)


def resolve_kernel(kernel, n_words: int) -> int:
    """Kernel id AUTO / the process default resolves to for rows of ``n_words`` (``STORM_b200_resolve_kernel``)."""
    return int(_lib.load().STORM_b200_resolve_kernel(_kernel_id(kernel), n_words))


def tile_rect(n_rows: int, tile: int, kernel=KERNEL_AUTO):
    """``STORM_b200_tile_rect`` (host only): rows (i0, i1, j0, j1) a tile covers."""
    L = _lib.load()
    v = [C.c_uint64() for _ in range(4)]
    _lib.check(L.STORM_b200_tile_rect(n_rows, _kernel_id(kernel), tile, *[C.byref(x) for x in v]), "STORM_b200_tile_rect")
    return tuple(int(x.value) for x in v)


def microbench(kind: int):
    L = _lib.load()
    rate, mhz = C.c_double(), C.c_double()
    _lib.check(L.STORM_b200_microbench(kind, C.byref(rate), C.byref(mhz)), "STORM_b200_microbench")
    return rate.value, mhz.value


def launch_count() -> int:
    return int(_lib.load().STORM_b200_launch_count())


def set_default_kernel(kernel) -> int:
    return _lib.load().STORM_b200_set_default_kernel(_kernel_id(kernel))


def set_umma_cta_group(cg: int) -> int:
    return _lib.load().STORM_b200_set_umma_cta_group(int(cg))


def set_storm_route(route) -> int:
    """``STORM_b200_set_storm_route``: 'auto' | 'sparse' | 'dense' | 'split' for whole-container STORM_t queries."""
    r = {"auto": 0, "sparse": 1, "dense": 2, "split": 3}[route] if isinstance(route, str) else int(route)
    return _lib.load().STORM_b200_set_storm_route(r)


def storm_split_model(n_rows: int, n_heavy: int, light_nnz: float, total_nnz: float, max_blocks: float, n_bitmap_blocks: float) -> float:
    """``STORM_b200_storm_split_model``: expected seconds of the split route (no device needed)."""
    return _lib.load().STORM_b200_storm_split_model(n_rows, n_heavy, light_nnz, total_nnz, max_blocks, n_bitmap_blocks)


LIST_ROUTES = {"auto": 0, "tile": 1, "probe": 2, "stream": 3}


def set_contig_list_route(route) -> int:
    """``STORM_b200_set_contig_list_route``: 'auto' | 'tile' | 'probe' | 'stream' for the contiguous *_list queries."""
    r = LIST_ROUTES[route] if isinstance(route, str) else int(route)
    return _lib.load().STORM_b200_set_contig_list_route(r)


def storm_route_model(n_rows: int, n_words: int, avg_nnz: float, avg_blocks: float, max_row_nnz: int, n_bitmap_blocks: int = 0,
                      fp4: bool = True, dense_resident: bool = True) -> dict:
    """``STORM_b200_storm_route_model``: expected seconds of a STORM_t query on either route (no device needed)."""
    out = (C.c_double * 2)()
    _lib.check(_lib.load().STORM_b200_storm_route_model(n_rows, n_words, avg_nnz, avg_blocks, max_row_nnz, n_bitmap_blocks,
                                                        int(fp4), int(dense_resident), out), "STORM_b200_storm_route_model")
    return {"dense_s": out[0], "sparse_s": out[1], "route": "dense" if out[0] < out[1] else "sparse"}


def set_storm_band_rows(rows: int) -> int:
    """``STORM_b200_set_storm_band_rows``: force the banded dense form of STORM_t queries (0 = size rule)."""
    return int(_lib.load().STORM_b200_set_storm_band_rows(int(rows)))


def set_sparse_flat(mode) -> int:
    """``STORM_b200_set_sparse_flat``: 'block' (0) | 'flat' (1) | 'stream' (2, default) kernels of the sparse route
    for containers without bitmap blocks; returns the previous mode."""
    m = {"block": 0, "flat": 1, "stream": 2}[mode] if isinstance(mode, str) else int(mode)
    return _lib.load().STORM_b200_set_sparse_flat(m)


def set_umma_variant(variant: int) -> int:
    """Code variant of the UMMA kernel (bit 0: suspended waits, bit 1: scaled expansion)."""
    return _lib.load().STORM_b200_set_umma_variant(int(variant))


def set_umma_wave_sync(on: bool) -> int:
    """Wave-synchronous tile schedule of the UMMA kernel (default on); returns the previous value."""
    return _lib.load().STORM_b200_set_umma_wave_sync(int(bool(on)))


def set_umma_reserved_sms(n: int) -> int:
    """SMs the persistent UMMA kernel leaves free (for a collective running beside it); returns the previous value."""
    return _lib.load().STORM_b200_set_umma_reserved_sms(int(n))


def set_umma_stream_k(on: bool) -> int:
    """Stream-K split of total-only UMMA queries with few tiles per SM (default on); returns the previous value."""
    return _lib.load().STORM_b200_set_umma_stream_k(int(bool(on)))


def set_umma_chain(on: bool) -> int:
    """Accumulator chaining of total-only UMMA queries (one drain per run of interior tiles, default on); returns the previous value."""
    return _lib.load().STORM_b200_set_umma_chain(int(bool(on)))


def pairw_devices(rows_per_device, n_words: Optional[int] = None, kernel=KERNEL_AUTO) -> int:
    """``STORM_b200_pairw_devices``: the same matrix resident on several devices (one CUDA tensor per device), tile raster
    sharded over them behind one C call, totals added on the host."""
    L = _lib.load()
    n = len(rows_per_device)
    ptrs = (C.c_void_p * n)(*[r.data_ptr() for r in rows_per_device])
    ids = (C.c_int * n)(*[r.device.index for r in rows_per_device])
    _, n_rows, stride = _rows_args(rows_per_device[0])
    return _query(L.STORM_b200_pairw_devices(ptrs, ids, n, n_rows, n_words or rows_per_device[0].shape[1], stride, _kernel_id(kernel)),
                  "STORM_b200_pairw_devices")


def set_clock_probe(on: bool) -> int:
    """``STORM_b200_set_clock_probe``: tensor-kernel launches record their clock64 / globaltimer deltas."""
    return _lib.load().STORM_b200_set_clock_probe(int(bool(on)))


def last_kernel_clock() -> dict:
    """``STORM_b200_last_kernel_clock``: clock64 ticks per microsecond of the last probed launch (mean, min, max over CTAs)."""
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    _lib.check(_lib.load().STORM_b200_last_kernel_clock(C.byref(a), C.byref(b), C.byref(c)), "STORM_b200_last_kernel_clock")
    return {"mhz": a.value, "min_mhz": b.value, "max_mhz": c.value}


def set_devices(n: int) -> int:
    """``STORM_b200_set_devices``: 0 = every visible device, n >= 1 = devices 0 .. n-1 for containers created and
    wrapper calls made afterwards; returns the previous count."""
    return _lib.load().STORM_b200_set_devices(int(n))


def set_device_list(ids) -> None:
    """``STORM_b200_set_device_list``: explicit ordinals (one may repeat: replicas on one device); () = the default."""
    ids = list(ids)
    arr = (C.c_int * max(1, len(ids)))(*ids)
    _lib.check(_lib.load().STORM_b200_set_device_list(arr if ids else None, len(ids)), "STORM_b200_set_device_list")


def get_devices() -> list:
    out = (C.c_int * 64)()
    n = _lib.load().STORM_b200_get_devices(out, 64)
    if n < 0:
        raise StormError(f"STORM_b200_get_devices failed ({n}): {_lib.last_error()}")
    return [out[i] for i in range(min(n, 64))]


def set_device_threads(on: bool) -> bool:
    """``STORM_b200_set_device_threads``: per-device calls of a multi-device query from one host thread per device (default) or from the caller alone."""
    return bool(_lib.load().STORM_b200_set_device_threads(int(bool(on))))


def last_error() -> str:
    return _lib.last_error()


def device_info(dev: int = 0) -> dict:
    L = _lib.load()
    name = C.create_string_buffer(128)
    sms, cc = C.c_int(), C.c_int()
    _lib.check(L.STORM_b200_device_info(dev, name, 128, C.byref(sms), C.byref(cc)), "STORM_b200_device_info")
    return {"name": name.value.decode(), "sm_count": sms.value, "cc": cc.value}
