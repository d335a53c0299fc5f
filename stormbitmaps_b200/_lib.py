"""ctypes binding of libstorm_b200.so (the C-ABI declared in include/*.h).

There is no fallback: if the shared object is missing the import fails with
instructions to build it, and if it loads but no sm_100 device is usable every
query raises :class:`StormError` carrying ``STORM_b200_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libstorm_b200.so")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
UINT64_MAX = 2**64 - 1

KERNEL_AUTO, KERNEL_POPC, KERNEL_UMMA, KERNEL_CSA, KERNEL_FP4 = 0, 1, 2, 3, 4
KERNEL_NAMES = {"auto": 0, "popc": 1, "umma": 2, "csa": 3, "fp4": 4, "b1": 5}


class StormError(RuntimeError):
    pass


# every symbol include/storm.h and include/storm_b200.h declare: (restype, argtypes)
SIGNATURES = {
    # ---- storm.h: dense model
    "STORM_contig_new": (C.c_void_p, [C.c_size_t]),
    "STORM_contig_free": (None, [C.c_void_p]),
    "STORM_contig_add": (C.c_int, [C.c_void_p, u32p, C.c_uint32]),
    "STORM_contig_clear": (C.c_int, [C.c_void_p]),
    "STORM_contig_pairw_intersect_cardinality": (C.c_uint64, [C.c_void_p]),
    "STORM_contig_pairw_intersect_cardinality_blocked": (C.c_uint64, [C.c_void_p, C.c_uint32]),
    "STORM_contig_pairw_intersect_cardinality_list": (C.c_uint64, [C.c_void_p]),
    "STORM_contig_pairw_intersect_cardinality_blocked_list": (C.c_uint64, [C.c_void_p, C.c_uint32]),
    # ---- storm.h: sparse model
    "STORM_new": (C.c_void_p, []),
    "STORM_free": (None, [C.c_void_p]),
    "STORM_add": (C.c_int, [C.c_void_p, u32p, C.c_uint32]),
    "STORM_clear": (C.c_int, [C.c_void_p]),
    "STORM_pairw_intersect_cardinality": (C.c_uint64, [C.c_void_p]),
    "STORM_pairw_intersect_cardinality_blocked": (C.c_uint64, [C.c_void_p, C.c_uint32]),
    "STORM_intersect_cardinality_square": (C.c_uint64, [C.c_void_p, C.c_void_p]),
    "STORM_serialized_size": (C.c_uint64, [C.c_void_p]),
    "STORM_bitmap_cont_new": (C.c_void_p, []),
    "STORM_bitmap_cont_init": (None, [C.c_void_p]),
    "STORM_bitmap_cont_free": (None, [C.c_void_p]),
    "STORM_bitmap_cont_add": (C.c_int, [C.c_void_p, u32p, C.c_uint32]),
    "STORM_bitmap_cont_clear": (C.c_int, [C.c_void_p]),
    "STORM_bitmap_cont_serialized_size": (C.c_uint32, [C.c_void_p]),
    "STORM_bitmap_new": (C.c_void_p, []),
    "STORM_bitmap_init": (None, [C.c_void_p]),
    "STORM_bitmap_free": (None, [C.c_void_p]),
    "STORM_bitmap_add": (C.c_int, [C.c_void_p, u32p, C.c_uint32]),
    "STORM_bitmap_add_scalar_only": (C.c_int, [C.c_void_p, u32p, C.c_uint32]),
    "STORM_bitmap_clear": (C.c_int, [C.c_void_p]),
    "STORM_bitmap_serialized_size": (C.c_uint32, [C.c_void_p]),
    # ---- storm.h: raw-buffer wrappers
    "STORM_wrapper_diag": (C.c_uint64, [C.c_uint32, u64p, C.c_uint32, C.c_void_p]),
    "STORM_wrapper_diag_blocked": (C.c_uint64, [C.c_uint32, u64p, C.c_uint32, C.c_void_p, C.c_uint32]),
    "STORM_wrapper_square": (C.c_uint64, [C.c_uint32, u64p, C.c_uint32, u64p, C.c_uint32, C.c_void_p]),
    "STORM_wrapper_diag_list": (C.c_uint64, [C.c_uint32, u64p, C.c_uint32, u32p, u32p, u32p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "STORM_wrapper_diag_list_blocked": (C.c_uint64, [C.c_uint32, u64p, C.c_uint32, u32p, u32p, u32p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]),
    "STORM_get_intersect_count_func": (C.c_void_p, [C.c_size_t]),
    "STORM_get_union_count_func": (C.c_void_p, [C.c_size_t]),
    "STORM_get_diff_count_func": (C.c_void_p, [C.c_size_t]),
    # ---- storm_b200.h
    "STORM_b200_row_popcounts_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]),
    "STORM_b200_pairw_op_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "STORM_b200_pairw_rect_op_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "STORM_b200_host_intersect_count": (C.c_uint64, [u64p, u64p, C.c_size_t]),
    "STORM_b200_host_union_count": (C.c_uint64, [u64p, u64p, C.c_size_t]),
    "STORM_b200_host_diff_count": (C.c_uint64, [u64p, u64p, C.c_size_t]),
    "STORM_b200_last_error": (C.c_char_p, []),
    "STORM_b200_version": (C.c_char_p, []),
    "STORM_b200_device_count": (C.c_int, []),
    "STORM_b200_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "STORM_b200_set_default_kernel": (C.c_int, [C.c_int]),
    "STORM_b200_pairw_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]),
    "STORM_b200_pairw_rect_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "STORM_b200_square_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "STORM_b200_resolve_kernel": (C.c_int, [C.c_int, C.c_uint32]),
    "STORM_b200_wrapper_diag_shard": (C.c_uint64, [C.c_uint64, u64p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]),
    "STORM_b200_pairw_tiles_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]),
    "STORM_b200_tiles_below_row": (C.c_int, [C.c_uint64, C.c_int, C.c_uint64, u64p, u64p]),
    "STORM_b200_tile_count": (C.c_uint64, [C.c_uint64, C.c_int, u32p, u32p]),
    "STORM_b200_shard_tiles": (C.c_int, [C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, u64p, u64p]),
    "STORM_b200_tile_rect": (C.c_int, [C.c_uint64, C.c_int, C.c_uint64, u64p, u64p, u64p, u64p]),
    "STORM_b200_contig_pairw_shard": (C.c_uint64, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]),
    "STORM_b200_contig_pairw_rect": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, u32p]),
    "STORM_b200_contig_device_rows": (C.c_void_p, [C.c_void_p, u64p]),
    "STORM_b200_contig_invalidate_device": (C.c_int, [C.c_void_p]),
    "STORM_b200_contig_add_bulk": (C.c_int, [C.c_void_p, u32p, u64p, C.c_uint64]),
    "STORM_b200_contig_last_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "STORM_b200_storm_pairw_shard": (C.c_uint64, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "STORM_b200_set_storm_route": (C.c_int, [C.c_int]),
    "STORM_b200_storm_split_model": (C.c_double, [C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double]),
    "STORM_b200_set_sparse_flat": (C.c_int, [C.c_int]),
    "STORM_b200_set_storm_band_rows": (C.c_uint64, [C.c_uint64]),
    "STORM_b200_storm_route_model": (C.c_int, [C.c_uint64, C.c_uint32, C.c_double, C.c_double, C.c_uint32, C.c_uint64, C.c_int, C.c_int,
                                              C.POINTER(C.c_double)]),
    "STORM_b200_set_contig_list_route": (C.c_int, [C.c_int]),
    "STORM_b200_contig_last_list_route": (C.c_int, [C.c_void_p]),
    "STORM_b200_storm_last_route": (C.c_int, [C.c_void_p]),
    "STORM_b200_storm_pairw_rect": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, u32p]),
    "STORM_b200_synth_uniform_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]),
    "STORM_b200_synth_geno_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]),
    "STORM_b200_microbench": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "STORM_b200_fp4_probe": (C.c_int, [u32p, C.c_uint32, C.POINTER(C.c_float)]),
    "STORM_b200_fp4_selftest": (C.c_int, []),
    "STORM_b200_fp4_probe_random": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]),
    "STORM_b200_set_umma_cta_group": (C.c_int, [C.c_int]),
    "STORM_b200_set_umma_variant": (C.c_int, [C.c_int]),
    "STORM_b200_set_umma_wave_sync": (C.c_int, [C.c_int]),
    "STORM_b200_set_umma_stream_k": (C.c_int, [C.c_int]),
    "STORM_b200_set_umma_chain": (C.c_int, [C.c_int]),
    "STORM_b200_set_umma_reserved_sms": (C.c_int, [C.c_int]),
    "STORM_b200_launch_count": (C.c_uint64, []),
    "STORM_b200_set_umma_l2_hints": (C.c_int, [C.c_int]),
    "STORM_b200_set_clock_probe": (C.c_int, [C.c_int]),
    "STORM_b200_last_kernel_clock": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "STORM_b200_pairw_tiles_device_ex": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "STORM_b200_set_devices": (C.c_int, [C.c_int]),
    "STORM_b200_set_device_list": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "STORM_b200_get_devices": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "STORM_b200_set_device_threads": (C.c_int, [C.c_int]),
    "STORM_b200_selftest_device_threads": (C.c_int, [C.c_int, C.c_int]),
    "STORM_b200_contig_device_count": (C.c_int, [C.c_void_p]),
    "STORM_b200_contig_add_dense": (C.c_int, [C.c_void_p, u64p, C.c_uint64, C.c_uint64]),
    "STORM_b200_contig_rehome": (C.c_int, [C.c_void_p]),
    "STORM_b200_pairw_devices": (C.c_uint64, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]),
    # ---- storm.h: per-pair host helpers
    "STORM_bitmap_add_with_scalar": (C.c_int, [C.c_void_p, u32p, C.c_uint32]),
    "STORM_intersect_vector16_cardinality": (C.c_uint64, [C.POINTER(C.c_uint16), C.POINTER(C.c_uint16), C.c_uint32, C.c_uint32]),
    "STORM_intersect_vector32_unsafe": (C.c_uint64, [u32p, u32p, C.c_uint32, C.c_uint32, u32p]),
    "STORM_intersect_bitmaps_scalar_list": (C.c_uint64, [u64p, u64p, u32p, u32p, C.c_uint32, C.c_uint32]),
    "STORM_bitmap_intersect_cardinality": (C.c_uint64, [C.c_void_p, C.c_void_p]),
    "STORM_bitmap_intersect_cardinality_func": (C.c_uint64, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "STORM_bitmap_cont_intersect_cardinality": (C.c_uint64, [C.c_void_p, C.c_void_p]),
    "STORM_bitmap_cont_intersect_cardinality_premade": (C.c_uint64, [C.c_void_p, C.c_void_p, C.c_void_p, u32p]),
}

_lib = None


def load() -> C.CDLL:
    """Load libstorm_b200.so and bind every declared symbol (raises if any is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m stormbitmaps_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def last_error() -> str:
    return load().STORM_b200_last_error().decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise StormError(f"{what} failed (code {rc}): {last_error()}")
