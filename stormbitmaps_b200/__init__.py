"""storm-b200: the StormBitmaps all-vs-all intersection-cardinality path on B200.

The product is ``libstorm_b200.so`` (hand-written sm_100a CUDA behind the
reference's C API, see include/storm.h and include/storm_b200.h); this package is
the thin Python mirror of that API plus the in-tree build script.
"""
from ._lib import KERNEL_AUTO, KERNEL_CSA, KERNEL_FP4, KERNEL_POPC, KERNEL_UMMA, StormError, load  # noqa: F401
from .api import (Storm, StormContiguous, alloc_rows, device_info, launch_count, microbench,  # noqa: F401
                  pairw_device, pairw_tiles_device, tiles_below_row, resolve_kernel, pairw_rect_device, pairw_op_device, pairw_rect_op_device, row_popcounts_device, set_default_kernel, set_umma_cta_group, set_umma_variant, set_umma_wave_sync, set_umma_stream_k, set_umma_chain, set_umma_reserved_sms, set_storm_route, set_sparse_flat, set_contig_list_route, storm_route_model, storm_split_model, shard_tiles, tile_rect, square_device, synth_geno_device,
                  synth_uniform_device, tile_count, wrapper_diag, wrapper_diag_ptr, wrapper_diag_shard_ptr,
                  resolved_kernel_name, wrapper_square, wrapper_diag_list, set_devices, set_device_list, get_devices, set_device_threads, last_error, set_clock_probe, last_kernel_clock, pairw_devices, set_storm_band_rows)

__all__ = ["Storm", "StormContiguous", "StormError", "load", "pairw_device", "pairw_tiles_device", "tiles_below_row", "resolve_kernel", "pairw_rect_device", "pairw_op_device", "pairw_rect_op_device", "row_popcounts_device", "square_device",
           "alloc_rows", "synth_uniform_device", "synth_geno_device", "wrapper_diag", "wrapper_diag_ptr",
           "wrapper_square", "wrapper_diag_shard_ptr", "resolved_kernel_name", "tile_count", "microbench", "launch_count", "set_default_kernel", "set_umma_cta_group", "set_umma_variant", "set_umma_wave_sync", "set_umma_stream_k", "set_umma_chain", "set_umma_reserved_sms", "set_storm_route", "set_sparse_flat", "set_contig_list_route", "storm_route_model", "storm_split_model", "shard_tiles", "tile_rect", "device_info", "wrapper_diag_list", "set_devices", "set_device_list", "get_devices", "set_device_threads", "last_error", "set_clock_probe", "last_kernel_clock", "pairw_devices", "set_storm_band_rows",
           "KERNEL_AUTO", "KERNEL_POPC", "KERNEL_UMMA", "KERNEL_CSA", "KERNEL_FP4"]
