"""ctypes front-end for the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Two shared objects are wrapped here (both built by ``oracle/Makefile``):

* ``oracle/_build/liboracle.so`` -- our C restatement (``storm_oracle.c``);
* ``oracle/_ref/libstorm_ref.so`` -- the UNMODIFIED reference (``storm.c`` +
  ``libalgebra.h`` from /root/reference) plus ``ref_shim.c``.  It is built in the
  development container and travels to the GPU box as a prebuilt file; nothing
  here reads /root/reference at run time.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may
import this module.  The product package ``stormbitmaps_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libstorm_ref.so")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u16p = C.POINTER(C.c_uint16)


def build(reference: str = "/root/reference") -> None:
    """Compile the restatement, and the reference when its sources are present."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isfile(os.path.join(reference, "storm.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", f"REF={reference}"])


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


def _u64(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


# --------------------------------------------------------------------------- #
# restatement
# --------------------------------------------------------------------------- #
class Oracle:
    def __init__(self, path: str = ORACLE_SO):
        if not os.path.isfile(path):
            build()
        self.lib = L = C.CDLL(path)
        sig = {
            "orc_intersect_count": (C.c_uint64, [u64p, u64p, C.c_size_t]),
            "orc_wrapper_diag": (C.c_uint64, [C.c_uint64, u64p, C.c_uint64]),
            "orc_rect_total": (C.c_uint64, [u64p, C.c_uint64] + [C.c_uint64] * 4),
            "orc_rect_counts": (None, [u64p, C.c_uint64] + [C.c_uint64] * 4 + [u32p]),
            "orc_wrapper_square": (C.c_uint64, [C.c_uint64, u64p, C.c_uint64, u64p, C.c_uint64]),
            "orc_wrapper_diag_op": (C.c_uint64, [C.c_uint64, u64p, C.c_uint64, C.c_int]),
            "orc_rect_counts_op": (None, [u64p, C.c_uint64] + [C.c_uint64] * 4 + [C.c_int, u32p]),
            "orc_wrapper_square_op": (C.c_uint64, [C.c_uint64, u64p, C.c_uint64, u64p, C.c_uint64, C.c_int]),
            "orc_colcount_total": (C.c_uint64, [C.c_uint64, u64p, C.c_uint64]),
            "orc_colcount_rect": (C.c_uint64, [u64p, C.c_uint64] + [C.c_uint64] * 4),
            "orc_contig_new": (C.c_void_p, [C.c_size_t]),
            "orc_contig_free": (None, [C.c_void_p]),
            "orc_contig_add": (C.c_int, [C.c_void_p, u32p, C.c_uint32]),
            "orc_contig_clear": (C.c_int, [C.c_void_p]),
            "orc_contig_pairw": (C.c_uint64, [C.c_void_p]),
            "orc_contig_pairw_blocked": (C.c_uint64, [C.c_void_p, C.c_uint32]),
            "orc_contig_pairw_list": (C.c_uint64, [C.c_void_p]),
            "orc_contig_pairw_blocked_list": (C.c_uint64, [C.c_void_p, C.c_uint32]),
            "orc_storm_new": (C.c_void_p, []),
            "orc_storm_free": (None, [C.c_void_p]),
            "orc_storm_add": (C.c_int, [C.c_void_p, u32p, C.c_uint32]),
            "orc_storm_clear": (C.c_int, [C.c_void_p]),
            "orc_storm_pairw": (C.c_uint64, [C.c_void_p, C.c_int]),
            "orc_storm_pairw_blocked": (C.c_uint64, [C.c_void_p, C.c_uint32, C.c_int]),
            "orc_storm_serialized_size": (C.c_uint64, [C.c_void_p]),
            "orc_storm_auto_bsize": (C.c_uint32, [C.c_void_p]),
            "orc_intersect_u16": (C.c_uint64, [u16p, u16p, C.c_uint32, C.c_uint32]),
            "orc_splitmix64": (C.c_uint64, [C.c_uint64]),
            "orc_draw_position": (C.c_uint32, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]),
            "orc_gen_row_positions": (C.c_uint32, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, u32p]),
            "orc_gen_dense_uniform": (None, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, u64p, C.c_uint64]),
            "orc_geno_threshold": (C.c_uint32, [C.c_uint64, C.c_uint64]),
            "orc_gen_dense_geno": (None, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, u64p, C.c_uint64]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args

    # ---- raw buffers ---------------------------------------------------- #
    def pair_count(self, a, b) -> int:
        a, b = _u64(a), _u64(b)
        return int(self.lib.orc_intersect_count(_ptr(a, u64p), _ptr(b, u64p), a.size))

    def wrapper_diag(self, vals: np.ndarray) -> int:
        vals = _u64(vals)
        n, w = vals.shape
        return int(self.lib.orc_wrapper_diag(n, _ptr(vals, u64p), w))

    def rect_total(self, vals, i0, i1, j0, j1) -> int:
        vals = _u64(vals)
        return int(self.lib.orc_rect_total(_ptr(vals, u64p), vals.shape[1], i0, i1, j0, j1))

    def rect_counts(self, vals, i0, i1, j0, j1) -> np.ndarray:
        vals = _u64(vals)
        out = np.zeros((i1 - i0, j1 - j0), dtype=np.uint32)
        self.lib.orc_rect_counts(_ptr(vals, u64p), vals.shape[1], i0, i1, j0, j1, _ptr(out, u32p))
        return out

    # set operations: op 0 intersect, 1 union, 2 diff (libalgebra.h:2985-3008 under storm.c:132-150)
    def wrapper_diag_op(self, vals, op: int) -> int:
        vals = _u64(vals)
        return int(self.lib.orc_wrapper_diag_op(vals.shape[0], _ptr(vals, u64p), vals.shape[1], op))

    def rect_counts_op(self, vals, i0, i1, j0, j1, op: int) -> np.ndarray:
        vals = _u64(vals)
        out = np.zeros((i1 - i0, j1 - j0), dtype=np.uint32)
        self.lib.orc_rect_counts_op(_ptr(vals, u64p), vals.shape[1], i0, i1, j0, j1, op, _ptr(out, u32p))
        return out

    def wrapper_square_op(self, v1, v2, op: int) -> int:
        v1, v2 = _u64(v1), _u64(v2)
        return int(self.lib.orc_wrapper_square_op(v1.shape[0], _ptr(v1, u64p), v2.shape[0], _ptr(v2, u64p), v1.shape[1], op))

    def wrapper_square(self, v1, v2) -> int:
        v1, v2 = _u64(v1), _u64(v2)
        return int(self.lib.orc_wrapper_square(v1.shape[0], _ptr(v1, u64p), v2.shape[0], _ptr(v2, u64p), v1.shape[1]))

    def colcount_total(self, vals) -> int:
        vals = _u64(vals)
        return int(self.lib.orc_colcount_total(vals.shape[0], _ptr(vals, u64p), vals.shape[1]))

    def colcount_rect(self, vals, i0, i1, j0, j1) -> int:
        vals = _u64(vals)
        return int(self.lib.orc_colcount_rect(_ptr(vals, u64p), vals.shape[1], i0, i1, j0, j1))

    # ---- generators ------------------------------------------------------ #
    def gen_row_positions(self, seed: int, row: int, n_draws: int, M: int) -> np.ndarray:
        out = np.empty(max(n_draws, 1), dtype=np.uint32)
        n = self.lib.orc_gen_row_positions(seed, row, n_draws, M, _ptr(out, u32p))
        return out[:n].copy()

    def gen_dense_uniform(self, seed: int, n_rows: int, n_draws: int, M: int, row0: int = 0,
                          n_words: Optional[int] = None) -> np.ndarray:
        w = n_words or (M + 63) // 64
        vals = np.zeros((n_rows, w), dtype=np.uint64)
        self.lib.orc_gen_dense_uniform(seed, row0, n_rows, n_draws, M, _ptr(vals, u64p), w)
        return vals

    def gen_dense_geno(self, seed: int, n_rows: int, M: int, row0: int = 0,
                       n_words: Optional[int] = None) -> np.ndarray:
        w = n_words or (M + 63) // 64
        vals = np.zeros((n_rows, w), dtype=np.uint64)
        self.lib.orc_gen_dense_geno(seed, row0, n_rows, M, _ptr(vals, u64p), w)
        return vals

    def intersect_u16(self, a, b) -> int:
        a = np.ascontiguousarray(a, dtype=np.uint16)
        b = np.ascontiguousarray(b, dtype=np.uint16)
        return int(self.lib.orc_intersect_u16(_ptr(a, u16p), _ptr(b, u16p), a.size, b.size))


class _Container:
    """Shared driver for the four container flavours (oracle/ref x contig/storm)."""

    def __init__(self, lib, handle, prefix_add, prefix_free):
        self.lib, self.h = lib, handle
        self._add, self._free = prefix_add, prefix_free

    def add(self, positions) -> int:
        p = np.ascontiguousarray(positions, dtype=np.uint32)
        return int(self._add(self.h, _ptr(p, u32p), p.size))

    def close(self):
        if self.h:
            self._free(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class OracleContig(_Container):
    def __init__(self, orc: Oracle, M: int):
        L = orc.lib
        super().__init__(L, L.orc_contig_new(M), L.orc_contig_add, L.orc_contig_free)

    def clear(self): return self.lib.orc_contig_clear(self.h)
    def pairw(self): return int(self.lib.orc_contig_pairw(self.h))
    def pairw_blocked(self, b): return int(self.lib.orc_contig_pairw_blocked(self.h, b))
    def pairw_list(self): return int(self.lib.orc_contig_pairw_list(self.h))
    def pairw_blocked_list(self, b): return int(self.lib.orc_contig_pairw_blocked_list(self.h, b))


class OracleStorm(_Container):
    def __init__(self, orc: Oracle):
        L = orc.lib
        super().__init__(L, L.orc_storm_new(), L.orc_storm_add, L.orc_storm_free)

    def clear(self): return self.lib.orc_storm_clear(self.h)
    def pairw(self, emulate_d1=False): return int(self.lib.orc_storm_pairw(self.h, int(emulate_d1)))
    def pairw_blocked(self, b=0, emulate_d1=False):
        return int(self.lib.orc_storm_pairw_blocked(self.h, b, int(emulate_d1)))
    def serialized_size(self): return int(self.lib.orc_storm_serialized_size(self.h))
    def auto_bsize(self): return int(self.lib.orc_storm_auto_bsize(self.h))


# --------------------------------------------------------------------------- #
# compiled reference
# --------------------------------------------------------------------------- #
def have_reference() -> bool:
    return os.path.isfile(REF_SO)


class Reference:
    """The unmodified reference (storm.h entry points + ref_shim.c helpers)."""

    def __init__(self, path: str = REF_SO):
        if not os.path.isfile(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = L = C.CDLL(path)
        sig = {
            "REF_intersect_count": (C.c_uint64, [u64p, u64p, C.c_size_t]),
            "REF_wrapper_diag": (C.c_uint64, [C.c_uint32, u64p, C.c_uint32]),
            "REF_wrapper_diag_blocked": (C.c_uint64, [C.c_uint32, u64p, C.c_uint32, C.c_uint32]),
            "REF_diag_rows": (C.c_uint64, [C.c_uint64, u64p, C.c_uint64, C.c_uint64, C.c_uint64]),
            "REF_rect_blocked": (C.c_uint64, [u64p, C.c_uint64] + [C.c_uint64] * 4 + [C.c_uint32]),
            "REF_wrapper_diag_op": (C.c_uint64, [C.c_uint32, u64p, C.c_uint32, C.c_int]),
            "REF_count_op": (C.c_uint64, [u64p, u64p, C.c_size_t, C.c_int]),
            "REF_cpuid": (C.c_int, []),
            "REF_kernel_name": (C.c_char_p, [C.c_size_t]),
            "REF_contig_scalar_cutoff": (C.c_uint32, [C.c_void_p]),
            "REF_contig_n_rows": (C.c_uint64, [C.c_void_p]),
            "REF_storm_n_rows": (C.c_uint32, [C.c_void_p]),
            "STORM_contig_new": (C.c_void_p, [C.c_size_t]),
            "STORM_contig_free": (None, [C.c_void_p]),
            "STORM_contig_add": (C.c_int, [C.c_void_p, u32p, C.c_uint32]),
            "STORM_contig_clear": (C.c_int, [C.c_void_p]),
            "STORM_contig_pairw_intersect_cardinality": (C.c_uint64, [C.c_void_p]),
            "STORM_contig_pairw_intersect_cardinality_blocked": (C.c_uint64, [C.c_void_p, C.c_uint32]),
            "STORM_contig_pairw_intersect_cardinality_list": (C.c_uint64, [C.c_void_p]),
            "STORM_contig_pairw_intersect_cardinality_blocked_list": (C.c_uint64, [C.c_void_p, C.c_uint32]),
            "STORM_new": (C.c_void_p, []),
            "STORM_free": (None, [C.c_void_p]),
            "STORM_add": (C.c_int, [C.c_void_p, u32p, C.c_uint32]),
            "STORM_clear": (C.c_int, [C.c_void_p]),
            "STORM_pairw_intersect_cardinality": (C.c_uint64, [C.c_void_p]),
            "STORM_pairw_intersect_cardinality_blocked": (C.c_uint64, [C.c_void_p, C.c_uint32]),
            "STORM_serialized_size": (C.c_uint64, [C.c_void_p]),
            "STORM_intersect_vector16_cardinality": (C.c_uint64, [u16p, u16p, C.c_uint32, C.c_uint32]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args

    def pair_count(self, a, b) -> int:
        a, b = _u64(a), _u64(b)
        return int(self.lib.REF_intersect_count(_ptr(a, u64p), _ptr(b, u64p), a.size))

    def wrapper_diag(self, vals) -> int:
        vals = _u64(vals)
        return int(self.lib.REF_wrapper_diag(vals.shape[0], _ptr(vals, u64p), vals.shape[1]))

    def wrapper_diag_op(self, vals, op: int) -> int:
        vals = _u64(vals)
        return int(self.lib.REF_wrapper_diag_op(vals.shape[0], _ptr(vals, u64p), vals.shape[1], op))

    def wrapper_diag_blocked(self, vals, bsize) -> int:
        vals = _u64(vals)
        return int(self.lib.REF_wrapper_diag_blocked(vals.shape[0], _ptr(vals, u64p), vals.shape[1], bsize))

    def diag_rows(self, vals, i0, i1) -> int:
        vals = _u64(vals)
        return int(self.lib.REF_diag_rows(vals.shape[0], _ptr(vals, u64p), vals.shape[1], i0, i1))

    def rect_blocked(self, vals, i0, i1, j0, j1, bsize) -> int:
        vals = _u64(vals)
        return int(self.lib.REF_rect_blocked(_ptr(vals, u64p), vals.shape[1], i0, i1, j0, j1, bsize))

    def kernel_name(self, n_words: int) -> str:
        return self.lib.REF_kernel_name(n_words).decode()

    def intersect_u16(self, a, b) -> int:
        a = np.ascontiguousarray(a, dtype=np.uint16)
        b = np.ascontiguousarray(b, dtype=np.uint16)
        return int(self.lib.STORM_intersect_vector16_cardinality(_ptr(a, u16p), _ptr(b, u16p), a.size, b.size))


class RefContig(_Container):
    def __init__(self, ref: Reference, M: int):
        L = ref.lib
        super().__init__(L, L.STORM_contig_new(M), L.STORM_contig_add, L.STORM_contig_free)

    def clear(self): return self.lib.STORM_contig_clear(self.h)
    def pairw(self): return int(self.lib.STORM_contig_pairw_intersect_cardinality(self.h))
    def pairw_blocked(self, b): return int(self.lib.STORM_contig_pairw_intersect_cardinality_blocked(self.h, b))
    def pairw_list(self): return int(self.lib.STORM_contig_pairw_intersect_cardinality_list(self.h))
    def pairw_blocked_list(self, b): return int(self.lib.STORM_contig_pairw_intersect_cardinality_blocked_list(self.h, b))
    def scalar_cutoff(self): return int(self.lib.REF_contig_scalar_cutoff(self.h))
    def n_rows(self): return int(self.lib.REF_contig_n_rows(self.h))


class RefStorm(_Container):
    def __init__(self, ref: Reference):
        L = ref.lib
        super().__init__(L, L.STORM_new(), L.STORM_add, L.STORM_free)

    def clear(self): return self.lib.STORM_clear(self.h)
    def pairw(self): return int(self.lib.STORM_pairw_intersect_cardinality(self.h))
    def pairw_blocked(self, b=0): return int(self.lib.STORM_pairw_intersect_cardinality_blocked(self.h, b))
    def serialized_size(self): return int(self.lib.STORM_serialized_size(self.h))
    def n_rows(self): return int(self.lib.REF_storm_n_rows(self.h))


# --------------------------------------------------------------------------- #
# numpy cross-check (third, independent implementation for tiny cases)
# --------------------------------------------------------------------------- #
def numpy_total(vals: np.ndarray) -> int:
    """sum_{i<j} popcount(X_i & X_j) via unpacked bits and an integer Gram matrix."""
    bits = np.unpackbits(np.ascontiguousarray(vals).view(np.uint8), axis=1, bitorder="little").astype(np.int64)
    g = bits @ bits.T
    return int(np.triu(g, 1).sum())


def numpy_counts(vals: np.ndarray) -> np.ndarray:
    bits = np.unpackbits(np.ascontiguousarray(vals).view(np.uint8), axis=1, bitorder="little").astype(np.int64)
    return np.triu(bits @ bits.T, 1).astype(np.uint32)


def positions_to_dense(rows, M: int) -> np.ndarray:
    w = (M + 63) // 64
    vals = np.zeros((len(rows), w), dtype=np.uint64)
    for r, p in enumerate(rows):
        p = np.asarray(p, dtype=np.uint64)
        np.bitwise_or.at(vals[r], (p >> np.uint64(6)).astype(np.int64), np.uint64(1) << (p & np.uint64(63)))
    return vals
