// common.cuh -- shared declarations for libstorm_b200.so (sm_100a only).
//
// The library computes the StormBitmaps pairwise intersection-cardinality path
// (reference: storm.c:1149-1241 dense, storm.c:877-961 sparse) on the GPU.
// Nothing in here falls back to the CPU: every launcher reports CUDA failures
// through set_error() and a negative return code.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "storm_b200.h"

namespace storm {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();
void count_launch(uint64_t n = 1);

#define STORM_CUDA_TRY(expr)                                                         \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            ::storm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                  \
            return STORM_B200_ECUDA;                                                 \
        }                                                                            \
    } while (0)

// ---- one dense job: a set of (row-block, column-block) tiles -----------------
//
// A and B are row-major bitmaps (64-bit words).  A tile (bi, bj) covers A rows
// [bi*TM, bi*TM+TM) x B rows [bj*TN, bj*TN+TN) where TM/TN belong to the kernel.
// Rectangle mode numbers tiles row-block-major.  Triangle mode (A == B, strict upper
// triangle) only has the tiles that intersect the triangle and numbers them in an
// L2-friendly raster: COLUMN blocks are taken in groups of TRI_GROUP; inside a group the
// walk is row-block-major (bi outer, bj inner).  Consecutive tile indices -- which is
// what the CTAs of one wave hold, and what a shard is a range of -- then share their B rows
// with the whole group and their A rows with TRI_GROUP - 1 neighbours, so a wave of 74 tiles
// touches ~18 distinct row blocks instead of 75.  Tile indices are
// also monotone in the largest row they touch: tiles of groups <= g only read rows below
// (g + 1) * TRI_GROUP * TN, which is what lets a host-buffer query start computing while the
// rest of the matrix is still being uploaded (contig.cu: wrapper_diag_impl).
// `group_prefix[g]` is the number of tiles in groups < g (n_groups + 1 entries, device).
// 12 column blocks per group: with the L2 hints of the tensor kernel (column blocks evict_last, row blocks evict_first)
// the column blocks of a group stay in L2 from wave to wave as long as they fit beside the streaming row blocks --
// 12 x 256 rows x 16 KiB = 50 MB on C3.  Same-box A/B (profiles/r02_ab_raster_group.jsonl): 8 -> 12 -> 16 blocks:
// C3 614.4 / 609.4 / 628.8 ms, 30 000 x 131 072 14.08 / 13.85 / 14.01 ms, 65 536 x 16 384 8.15 / 7.99 / 8.13 ms.
#ifndef STORM_TRI_GROUP
#define STORM_TRI_GROUP 12
#endif
constexpr uint32_t TRI_GROUP = STORM_TRI_GROUP;

struct DenseJob {
    const uint64_t* A;
    const uint64_t* B;
    uint64_t strideA, strideB;   // words
    uint64_t nA, nB;             // valid rows
    uint32_t n_words;            // words per row that carry data
    uint64_t i_off, j_off;       // global row index of A row 0 / B row 0 (for the j>i mask)
    int strict_upper;            // count / emit only pairs with global j > global i
    int triangle;                // 1: tile list is the triangle raster (needs group_prefix)
    const uint64_t* group_prefix;  // device, n_groups + 1 entries (triangle mode)
    uint32_t n_bi, n_bj;         // tile grid extents
    uint64_t tile_begin, tile_end;  // this launch handles tiles [begin, end)
    uint32_t* out;               // optional per-pair counts, out[(i)*ld + j] relative to A/B row 0
    uint64_t ld;
    unsigned long long* total;   // optional, accumulated with atomicAdd
    // Optional wave counter of the persistent UMMA kernel (zeroed before the launch): CTAs start the
    // loads of their next tile only when every CTA has issued the loads of the current one, so the
    // tiles of a wave walk K in step and share their row blocks in L2.  A performance hint only:
    // the wait is bounded and results never depend on it.
    unsigned int* wave_sync;
    // Set by the UMMA launcher for total-only jobs with few tiles per CTA: the persistent CTAs split the
    // (tile, K chunk) units evenly instead of whole tiles (a total is a sum, so any K split of a tile
    // is valid); removes the tail wave and keeps every SM busy when there are fewer tiles than SMs.
    uint32_t stream_k;
    // Set by the UMMA launcher for total-only jobs: up to chain_max consecutive interior segments of one
    // CTA (pair) accumulate into the same tensor-memory accumulator and are drained once (a total is a sum,
    // so the accumulator may hold the sum of several tiles as long as no element can overflow its exact
    // range); 0 or 1 = every segment is drained on its own.
    uint32_t chain_max;
    // SMs the persistent UMMA kernel leaves to work running beside it (a collective of a multi-process caller); -1 = the
    // process default (STORM_b200_set_umma_reserved_sms).  Per launch, so that no caller has to flip a global mid-query.
    int reserved_sms;
    // Optional clock probe (STORM_b200_set_clock_probe): per CTA {clock64 delta, %globaltimer delta in ns} around the
    // kernel's main loop, from which the host derives the SM clock the launch actually ran at (the board's power cap
    // lowers it below what nvidia-smi reports for long tensor launches).
    unsigned long long* clk;
    // Set by the UMMA launcher for per-pair jobs whose output matrix a tensor map can describe (16-byte aligned base, ld a
    // multiple of 4): the counts leave through TMA stores instead of per-thread stores.
    int out_tma;
    // Set by the UMMA launcher for triangle jobs: packed-row boxes of the column blocks (B side: the 8 blocks of a raster
    // group are shared by every wave of the group) are loaded with an L2 evict_last hint, those of the row blocks
    // (A side: new ones every wave) with evict_first, so that the former stay resident across waves.
    int l2_hints;
};

// Last row block of column block bj that intersects the strict upper triangle when A == B (square
// matrix, same origin): the block holding row (bj + 1) * TN - 2, clipped to the matrix.
__host__ __device__ inline uint32_t tri_iend(uint32_t bj, uint32_t n_bi, uint32_t TM, uint32_t TN) {
    const uint64_t last = (((uint64_t)bj + 1) * TN - 2) / TM;
    return last < n_bi ? (uint32_t)last : n_bi - 1;
}

// Number of tiles of the column-block group starting at column block c0 (host builds the prefix with it).
__host__ __device__ inline uint64_t tri_group_tiles(uint32_t c0, uint32_t n_bi, uint32_t n_bj, uint32_t TM, uint32_t TN) {
    uint64_t n = 0;
    for (uint32_t bj = c0; bj < c0 + TRI_GROUP && bj < n_bj; ++bj) n += (uint64_t)tri_iend(bj, n_bi, TM, TN) + 1;
    return n;
}

// Tile u of the group whose first column block is c0 -> (bi, bj): row-block-major inside the group.
__host__ __device__ inline void tri_group_coords(uint32_t c0, uint64_t u, uint32_t n_bi, uint32_t n_bj, uint32_t TM, uint32_t TN,
                                                 uint32_t& bi, uint32_t& bj) {
    const uint32_t cols = (c0 + TRI_GROUP <= n_bj) ? TRI_GROUP : n_bj - c0;      // column blocks in this group
    const uint32_t ifull = tri_iend(c0, n_bi, TM, TN);                           // up to here every column has a tile
    const uint64_t n_full = ((uint64_t)ifull + 1) * cols;
    // (a group has at most TRI_GROUP * n_bi < 2^32 tiles: 32-bit division)
    if (u < n_full) { bi = (uint32_t)u / cols; bj = c0 + (uint32_t)u % cols; return; }
    u -= n_full;
    for (uint32_t i = ifull + 1;; ++i) {                                         // ramp towards the diagonal: a suffix of the columns
        uint32_t c = 0;
        while (c < cols && tri_iend(c0 + c, n_bi, TM, TN) < i) ++c;
        const uint32_t valid = cols - c;
        if (u < valid || valid == 0) { bi = i; bj = c0 + c + (uint32_t)u; return; }   // (valid == 0 cannot happen for u < group size)
        u -= valid;
    }
}

// Map a linear tile index to (bi, bj).  `prefix` is job.group_prefix (device) or its host copy.
__host__ __device__ inline void tile_coords_tri(const uint64_t* prefix, uint32_t n_bi, uint32_t n_bj, uint64_t t, uint32_t TM, uint32_t TN,
                                                uint32_t& bi, uint32_t& bj) {
    const uint32_t n_groups = (n_bj + TRI_GROUP - 1) / TRI_GROUP;
    uint32_t lo = 0, hi = n_groups;          // largest g with prefix[g] <= t
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (prefix[mid] <= t) lo = mid; else hi = mid;
    }
    tri_group_coords(lo * TRI_GROUP, t - prefix[lo], n_bi, n_bj, TM, TN, bi, bj);
}

__device__ inline void tile_coords(const DenseJob& job, uint64_t t, uint32_t TM, uint32_t TN,
                                   uint32_t& bi, uint32_t& bj) {
    if (!job.triangle) {
        bi = (uint32_t)(t / job.n_bj);
        bj = (uint32_t)(t % job.n_bj);
        return;
    }
    tile_coords_tri(job.group_prefix, job.n_bi, job.n_bj, t, TM, TN, bi, bj);
}

// tile_coords with the last group's bounds kept in registers: a persistent CTA walks its tiles in increasing
// order, so most lookups stay inside the group of the previous one and need no global load at all.
struct TileCursor {
    uint64_t lo = 1, hi = 0;     // tiles [lo, hi) are group g (empty before the first lookup)
    uint32_t g = 0;
    __device__ __forceinline__ void coords(const DenseJob& job, uint64_t t, uint32_t TM, uint32_t TN, uint32_t& bi, uint32_t& bj) {
        if (!job.triangle) {
            bi = (uint32_t)(t / job.n_bj);
            bj = (uint32_t)(t % job.n_bj);
            return;
        }
        if (t < lo || t >= hi) {
            const uint32_t n_groups = (job.n_bj + TRI_GROUP - 1) / TRI_GROUP;
            uint32_t a = 0, b = n_groups;        // largest g with prefix[g] <= t
            while (b - a > 1) {
                const uint32_t mid = (a + b) >> 1;
                if (job.group_prefix[mid] <= t) a = mid; else b = mid;
            }
            g = a;
            lo = job.group_prefix[a];
            hi = job.group_prefix[a + 1];
        }
        tri_group_coords(g * TRI_GROUP, t - lo, job.n_bi, job.n_bj, TM, TN, bi, bj);
    }
};

// ---- launchers (one per kernel family) ---------------------------------------
struct TileShape { uint32_t tm, tn; };

TileShape popc_tile_shape();
int launch_dense_popc(const DenseJob& job, cudaStream_t stream);
TileShape csa_tile_shape();
int launch_dense_csa(const DenseJob& job, cudaStream_t stream);
TileShape b1_tile_shape();
int launch_dense_b1(const DenseJob& job, cudaStream_t stream);   // mma.sync .b1 AND + POPC (emulated on sm_100a; for the record)
TileShape umma_tile_shape();
TileShape umma_pairs_tile_shape();   // ... of a job with per-pair output (the tensor kernels' two-accumulator form)
int launch_dense_umma(const DenseJob& job, cudaStream_t stream);
// UMMA needs at least one full K step of 128 bits and 16-byte aligned rows.
bool umma_supports(const DenseJob& job);
bool umma_fp4_supports(const DenseJob& job);          // + every pair count below 2^24 (fp32-exact)
int launch_dense_fp4(const DenseJob& job, cudaStream_t stream);   // same kernel, kind::mxf4 form
// int8 ops per second of the UMMA kernel's own instruction issued back to back (cta_group 1 or 2).
int umma_peak_ops(int cg, double* ops_per_s, double* clock64_mhz);
int fp4_selftest_ok();                          // 1 if this device accumulates E2M1 bit products exactly (cached)
int fp4_peak_ops(int cg, double* ops_per_s, double* clock64_mhz, int n = 256);   // tcgen05.mma kind::mxf4 issue-rate probe (fp4_probe.cu)

int launch_synth_uniform(uint64_t* d_rows, uint64_t n_rows, uint64_t stride, uint32_t M,
                         uint32_t n_draws, uint64_t seed, uint64_t row0, cudaStream_t stream);
int launch_synth_geno(uint64_t* d_rows, uint64_t n_rows, uint64_t stride, uint32_t M,
                      uint64_t seed, uint64_t row0, cudaStream_t stream);
int launch_scatter_positions(uint64_t* d_rows, uint64_t stride, const uint32_t* d_pos,
                             const uint64_t* d_off, uint64_t n_rows, cudaStream_t stream);

// ---- small device helpers ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace storm
