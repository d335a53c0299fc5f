// runtime.h -- host-side helpers shared by the C-ABI translation units.
#pragma once

#include <vector>

#include "common.cuh"

namespace storm {

int require_device();
int default_kernel();
// cudaMemcpy host -> device that has reached the device on return (pageable sources are only staged when cudaMemcpy
// returns; the readers run on non-blocking streams, possibly from other host threads).
int copy_to_device_now(void* dst, const void* src, size_t bytes);

TileShape tile_shape_for(int kernel);
uint64_t triangle_prefix(uint64_t n_rows, TileShape ts, std::vector<uint64_t>* prefix, uint32_t* n_bi, uint32_t* n_bj);
void shard_range(uint64_t n_tiles, uint32_t shard, uint32_t n_shards, uint64_t* begin, uint64_t* end);
int resolve_kernel(int kernel, const DenseJob& job);
// Alignment / stride / width checks every dense entry point applies to a device matrix before any launch.
int check_rows(const uint64_t* d_rows, uint64_t stride, uint32_t n_words);
int launch_dense(int kernel, const DenseJob& job, cudaStream_t stream);

// Upper-triangle total of a device matrix (shard of the tile raster), accumulated into *d_total.
int pairw_triangle(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words, uint64_t stride,
                   uint32_t shard, uint32_t n_shards, int kernel, uint64_t* d_total, cudaStream_t stream);
// The same over an explicit range [tile_begin, tile_end) of the triangle raster.
int pairw_triangle_range(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words, uint64_t stride,
                         uint64_t tile_begin, uint64_t tile_end, int kernel, uint64_t* d_total, cudaStream_t stream,
                         int reserved_sms = -1);
// Kernel id AUTO resolves to for square-matrix rows of this geometry on the current device.
int resolve_kernel_for_rows(int kernel, const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words, uint64_t stride);
// All pairs of A rows x B rows (optionally only global j > i), counts and/or total.
int pairw_rect(const uint64_t* dA, uint64_t nA, uint64_t strideA, uint64_t i_off,
               const uint64_t* dB, uint64_t nB, uint64_t strideB, uint64_t j_off,
               uint32_t n_words, int strict_upper, int kernel,
               uint32_t* d_out, uint64_t ld, uint64_t* d_total, cudaStream_t stream);

// The same rectangle (or whole triangle) under a set operation STORM_B200_OP_* (setops.cu).
int pairw_rect_op(const uint64_t* dA, uint64_t nA, uint64_t strideA, uint64_t i_off,
                  const uint64_t* dB, uint64_t nB, uint64_t strideB, uint64_t j_off,
                  uint32_t n_words, int strict_upper, int op, int kernel, bool triangle,
                  uint32_t* d_out, uint64_t ld, uint64_t* d_total, cudaStream_t stream);
// Which set operation a caller-supplied per-pair kernel pointer stands for (contig.cu).
int op_of_compute_func(const STORM_compute_func f);

// Sparse-row probe (storm.c:108-129) for the contiguous *_list entry points: pairs
// (s, x) where s walks `d_sparse_rows` and x every row that pairs with it once.
int launch_contig_probe(const uint64_t* d_rows, uint64_t stride, uint64_t n_rows,
                        const uint32_t* d_is_sparse, const uint32_t* d_sparse_rows, uint64_t n_sparse,
                        const uint32_t* d_pos, const uint64_t* d_pos_off,
                        unsigned long long* d_total, cudaStream_t stream);
// Gather rows idx[0..n) of src into a compact arena dst (same stride).
int launch_gather_rows(uint64_t* dst, const uint64_t* src, uint64_t stride, const uint32_t* d_idx, uint64_t n,
                       cudaStream_t stream);

// Row-group stream kernel (sparse.cu): totals over rows given as flat position lists (CSR offsets + absolute
// uint32 positions on the device).  stream_groups cuts the rows i into groups (false: some row is too long);
// stream_seconds is the fitted cost model the route decisions use.
bool stream_groups(const uint32_t* row_nnz, uint64_t n_rows, std::vector<uint32_t>* group_start);
double stream_seconds(double pairs, double avg_nnz, uint32_t ranges = 1);
uint32_t stream_ranges(double avg_nnz, uint32_t span, uint32_t* range_bits);
int launch_sparse_stream(const uint64_t* a_off, const uint32_t* a_pos, const std::vector<uint32_t>& h_group_start,
                         const uint32_t* d_group_start, const uint64_t* b_off, const uint32_t* b_pos, uint64_t b_total_nnz,
                         uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, int strict_upper,
                         uint32_t shard, uint32_t n_shards, unsigned long long* d_total, cudaStream_t stream);

}  // namespace storm
