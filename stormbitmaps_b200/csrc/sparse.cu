// sparse.cu -- STORM_t: the Roaring-like sparse model.
//
// Host side: the public containers of storm.h:158-178 rebuilt from scratch
// (reference builders: storm.c:398-569 blocks, :659-758 rows, :827-875 top).
// Device side: a flattened mirror (CSR of blocks + one u16 pool + one bitmap
// pool) and the pairwise kernel that replaces storm.c:877-961 with its per-block
// 4-way dispatch (storm.c:618-656) and block-id merge (storm.c:75-106, 790-814).
//
// Kernel shape: one CTA owns row i and a slice of partner
// rows j.  Unless row i is tiny, its blocks are expanded once into shared-memory
// bitmaps (8 KiB per block); each warp then walks partner rows, merges the two
// sorted block-id lists and, per shared block, dispatches on density:
//     j list            -> probe j's values into i's shared bitmap
//     j bitmap, i list (<= 256 values) -> probe i's values into j's global bitmap
//     j bitmap otherwise -> AND + POPC over the 1024 words
//     i tiny (<= 64 values in the row, never expanded):
//         j bitmap      -> probe i's values into j's bitmap
//         j list        -> sorted-list search intersection (the GPU form of the
//                          reference's merge, storm.c:4-73)
// All four give the exact |i AND j| for sorted unique inputs; the reference's
// probe defect D1 (storm.c:636,644) is not reproduced.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <new>
#include <vector>

#include "common.cuh"
#include "devices.h"
#include "runtime.h"

namespace storm {
namespace {

constexpr uint32_t BLOCK_BITS = STORM_DEFAULT_BLOCK_SIZE;          // 65536
constexpr uint32_t BLOCK_WORDS = BLOCK_BITS / 64;                  // 1024
constexpr uint32_t LIST_THRESHOLD = STORM_DEFAULT_SCALAR_THRESHOLD;  // 4096
constexpr uint32_t BITMAP_FLAG = 0x80000000u;
constexpr uint32_t ROW_HAS_BITMAP = 0x80000000u;                   // device copy of row_nnz: the row holds a bitmap block
constexpr uint32_t ROW_HEAVY = 0x40000000u;                        // ... the row is a heavy row of the split route (below)

// Threads per CTA: the kernel is latency-bound (dependent global loads of the merge, random probes), so it
// wants every warp it can get next to one CTA's shared bitmaps: 1024 threads when the shared-memory slots
// leave room for a single CTA per SM (8 warps gave 12.5 % occupancy and 17 % issue utilisation on a C4-like
// matrix, profiles/r01_sparse_c4_ncu_full.md), 512 below that (several CTAs per SM).
constexpr int SP_MAX_THREADS = 1024;
constexpr int SP_MAX_WARPS = SP_MAX_THREADS / 32;
constexpr size_t SP_ONE_CTA_SMEM = 100 * 1024;   // above this only one CTA fits on an SM
constexpr uint32_t SP_SLICE = 1024;        // partner rows per CTA
constexpr uint32_t SP_MAXB_CAP = 24;       // row-i blocks resident in shared memory per pass (24 x 8 KiB)
constexpr uint32_t SP_TINY_NNZ = 64;       // rows with <= this many values are never expanded
constexpr uint32_t SP_PROBE_I_MAX = 256;   // i-list probes into a j-bitmap up to this length

// Flattened device view of one STORM_t.
struct SparseView {
    const uint32_t* row_ptr;   // n_rows + 1: first block of each row
    const uint32_t* row_nnz;   // values per row; ROW_HAS_BITMAP set if any of the row's blocks is a bitmap block, ROW_HEAVY
                               // if it has one or holds more values than a row group of the stream kernel may
    const uint32_t* blk_id;    // block index (value / 65536), ascending within a row
    const uint32_t* blk_len;   // number of values; BITMAP_FLAG set for bitmap blocks
    const uint64_t* blk_off;   // list: element offset into `lists`; bitmap: word offset into `words`
    const uint16_t* lists;
    const uint64_t* words;
    uint32_t n_rows;
    // flat form (ensure_flat; NULL until built): every value of a row as an absolute position, CSR over rows
    const uint64_t* pos_off;   // n_rows + 1
    const uint32_t* pos;
    uint32_t n_blk_span;       // largest block index + 1: positions are below n_blk_span * 65536
    float avg_nnz;             // values per row on average (sub-warp width of the flat kernel)
};

struct SparseJob {
    SparseView A, B;
    uint64_t i0, i1, j0, j1;
    int strict_upper;          // same container: only pairs with j > i
    uint32_t shard, n_shards;  // rows i are dealt round-robin to shards
    uint32_t maxb;             // shared-memory slots (<= SP_MAXB_CAP)
    uint32_t slice;            // partner rows per CTA (SP_SLICE; fewer when there are few rows i)
    // split route: rows i are i_list[i0 .. i1) (the heavy rows of the container, ascending), the partner rows all of
    // [j0, j1), and pair (i, j) counts iff j is a light row or j > i -- every pair with a heavy row exactly once
    const uint32_t* i_list;
    uint32_t* out; uint64_t ld;
    unsigned long long* total;
};

__device__ __forceinline__ uint32_t probe_bit64(const uint64_t* words, uint32_t v) {
    return (uint32_t)((words[v >> 6] >> (v & 63)) & 1ull);
}

// |a-list ∩ b-list| contribution of one lane: values a[k], k = lane, lane+32, ... searched in sorted b.
__device__ __forceinline__ uint32_t search_intersect(const uint16_t* a, uint32_t na, const uint16_t* b, uint32_t nb, uint32_t lane) {
    uint32_t c = 0;
    for (uint32_t k = lane; k < na; k += 32) {
        const uint16_t v = a[k];
        uint32_t lo = 0, hi = nb;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (b[mid] < v) lo = mid + 1; else hi = mid;
        }
        c += (lo < nb && b[lo] == v);
    }
    return c;
}

// Set bits of the shared bitmap `ab` among the two block-relative values packed in `w`.
__device__ __forceinline__ uint32_t probe_pair(const uint32_t* ab, uint32_t w) {
    const uint32_t v0 = w & 0xFFFFu, v1 = w >> 16;
    return ((ab[v0 >> 5] >> (v0 & 31)) & 1u) + ((ab[v1 >> 5] >> (v1 & 31)) & 1u);
}

__global__ void __launch_bounds__(SP_MAX_THREADS, 1) sparse_pairs_kernel(const SparseJob job) {
    extern __shared__ __align__(16) uint32_t s_bits[];            // maxb x 2048 32-bit words
    __shared__ uint32_t s_id[SP_MAXB_CAP], s_len[SP_MAXB_CAP];
    __shared__ uint64_t s_off[SP_MAXB_CAP];
    __shared__ unsigned long long warp_part[SP_MAX_WARPS];
    const uint32_t SP_THREADS = blockDim.x, SP_WARPS = blockDim.x >> 5;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t i = job.i0 + job.shard + (uint64_t)blockIdx.x * job.n_shards;
    if (i >= job.i1) return;
    const bool heavy_rows = job.i_list != nullptr;
    if (heavy_rows) i = job.i_list[i];
    uint64_t jbeg = job.j0;
    if (job.strict_upper && jbeg < i + 1) jbeg = i + 1;
    const uint64_t js0 = jbeg + (uint64_t)blockIdx.y * job.slice;
    if (js0 >= job.j1) return;
    const uint64_t js1 = min(js0 + (uint64_t)job.slice, job.j1);

    const SparseView& A = job.A;
    const SparseView& B = job.B;
    const uint32_t rb = A.row_ptr[i], nbi = A.row_ptr[i + 1] - rb;
    if (nbi == 0) return;                                          // empty row: all counts 0 (out is pre-zeroed)
    // tiny rows are never expanded into shared memory: every block must be a list.  (A block built from >= 4096
    // values with duplicates is a bitmap block holding few bits -- storm.c:745-748 decides on the values handed in,
    // not on the bits set -- and carries ROW_HAS_BITMAP, so the compare below is false for its row.)
    const bool tiny = A.row_nnz[i] <= SP_TINY_NNZ;

    unsigned long long cta_lane_total = 0;
    for (uint32_t g = 0; g < nbi; g += job.maxb) {
        const uint32_t ng = min(job.maxb, nbi - g);
        __syncthreads();                                           // previous pass is done with shared memory
        if (tid < ng) {
            s_id[tid] = A.blk_id[rb + g + tid];
            s_len[tid] = A.blk_len[rb + g + tid];
            s_off[tid] = A.blk_off[rb + g + tid];
        }
        if (!tiny) {
            uint4* z = reinterpret_cast<uint4*>(s_bits);
            for (uint32_t k = tid; k < ng * 512; k += SP_THREADS) z[k] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        if (!tiny) {
            for (uint32_t s = 0; s < ng; ++s) {
                const uint32_t len = s_len[s];
                if (len & BITMAP_FLAG) {
                    const uint4* src = reinterpret_cast<const uint4*>(A.words + s_off[s]);
                    uint4* dst = reinterpret_cast<uint4*>(s_bits + s * 2048);
                    for (uint32_t k = tid; k < 512; k += SP_THREADS) dst[k] = src[k];
                } else {
                    const uint16_t* src = A.lists + s_off[s];
                    for (uint32_t k = tid; k < len; k += SP_THREADS) {
                        const uint32_t v = src[k];
                        atomicOr(&s_bits[s * 2048 + (v >> 5)], 1u << (v & 31));
                    }
                }
            }
            __syncthreads();
        }

        for (uint64_t j = js0 + warp; j < js1; j += SP_WARPS) {
            if (heavy_rows && (j == i || (j < i && (B.row_nnz[j] & ROW_HEAVY)))) continue;   // (warp-uniform)
            uint32_t c = 0;
            uint32_t y = B.row_ptr[j];
            const uint32_t ye = B.row_ptr[j + 1];
            uint32_t x = 0;
            while (x < ng && y < ye) {                             // storm.c:75-106: merge of sorted block ids
                const uint32_t ida = s_id[x], idb = B.blk_id[y];
                if (ida < idb) { ++x; continue; }
                if (ida > idb) { ++y; continue; }
                const uint32_t la = s_len[x], lb = B.blk_len[y];
                const uint32_t na = la & ~BITMAP_FLAG, nb = lb & ~BITMAP_FLAG;
                const uint64_t ob = B.blk_off[y];
                if (lb & BITMAP_FLAG) {
                    const uint64_t* bw = B.words + ob;
                    if (!(la & BITMAP_FLAG) && (tiny || na <= SP_PROBE_I_MAX)) {
                        const uint16_t* al = A.lists + s_off[x];   // list -> bitmap probe (storm.c:632-646, exact)
                        for (uint32_t k = lane; k < na; k += 32) c += probe_bit64(bw, al[k]);
                    } else {                                       // bitmap x bitmap (storm.c:648-650)
                        const uint4* b4 = reinterpret_cast<const uint4*>(bw);
                        const uint4* a4 = reinterpret_cast<const uint4*>(s_bits + x * 2048);
#pragma unroll 4
                        for (uint32_t k = lane; k < 512; k += 32) {
                            const uint4 p = b4[k], q = a4[k];
                            c += __popc(p.x & q.x) + __popc(p.y & q.y) + __popc(p.z & q.z) + __popc(p.w & q.w);
                        }
                    }
                } else {
                    const uint16_t* bl = B.lists + ob;
                    if (tiny) {                                    // list x list (storm.c:628-630)
                        c += search_intersect(A.lists + s_off[x], na, bl, nb, lane);
                    } else {                                       // j's values probed into i's shared bitmap
                        // lists start on 16-byte boundaries of the pool (sync_mirror): eight values per load
                        const uint32_t* ab = s_bits + x * 2048;
                        const uint4* bl4 = reinterpret_cast<const uint4*>(bl);
                        const uint32_t n8 = nb >> 3;
                        for (uint32_t k = lane; k < n8; k += 32) {
                            const uint4 q = __ldg(bl4 + k);
                            c += probe_pair(ab, q.x) + probe_pair(ab, q.y) + probe_pair(ab, q.z) + probe_pair(ab, q.w);
                        }
                        for (uint32_t k = (n8 << 3) + lane; k < nb; k += 32) {
                            const uint32_t v = bl[k];
                            c += (ab[v >> 5] >> (v & 31)) & 1u;
                        }
                    }
                }
                ++x; ++y;
            }
            cta_lane_total += c;
            if (job.out) {
                c = (uint32_t)warp_sum(c);
                if (lane == 0) {
                    uint32_t* o = job.out + (i - job.i0) * job.ld + (j - job.j0);
                    *o = (g == 0) ? c : *o + c;
                }
            }
        }
    }
    if (job.total) {
        const unsigned long long w = warp_sum(cta_lane_total);
        if (lane == 0) warp_part[warp] = w;
        __syncthreads();
        if (tid == 0) {
            unsigned long long t = 0;
            for (uint32_t k = 0; k < SP_WARPS; ++k) t += warp_part[k];
            if (t) atomicAdd(job.total, t);
        }
    }
}

// ---- flat probe kernel: the merge/probe path for rows that hold FEW values --------------------------------
// When no row has a bitmap block and a whole row fits shared memory as a bitmap (8 KiB per 65 536 bits), the
// per-pair block-id merge (storm.c:75-106) and the per-block dispatch (storm.c:618-656) collapse into one case:
// row i is expanded ONCE per CTA into a shared bitmap over the full row width and every value of a partner row
// j -- kept as an absolute position in a flat CSR mirror -- is probed into it (the list -> bitmap probe of
// storm.c:632-646 with the block arithmetic folded into the position).  A sub-warp of G lanes owns one partner
// row (G = 1 for rows of a few values: 32 pairs per warp step; G = 32 for hundreds), two dependent loads per pair
// (row offsets -> positions) instead of four per block.  The block kernel above spent ~12 000 cycles per pair
// on C2's 104-value level (dependent loads of the merge, one warp per pair); this one is bound by the probes.
constexpr size_t FLAT_MAX_SMEM = 160 * 1024;     // whole-row bitmap: rows up to 1 310 720 bits
constexpr uint32_t FLAT_SLICE = 4096;            // partner rows per CTA

// Flat row k = container row rows[k] (rows == NULL: row k), n of them.
__global__ void __launch_bounds__(256) flatten_rows_kernel(const SparseView v, const uint32_t* rows, uint32_t n, const uint64_t* pos_off, uint32_t* pos) {
    const uint32_t k_row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k_row >= n) return;
    const uint32_t row = rows ? rows[k_row] : k_row;
    uint64_t o = pos_off[k_row];
    for (uint32_t b = v.row_ptr[row]; b < v.row_ptr[row + 1]; ++b) {
        const uint32_t len = v.blk_len[b];                        // (no bitmap blocks: flat_eligible / light rows)
        const uint32_t base = v.blk_id[b] << 16;
        const uint16_t* src = v.lists + v.blk_off[b];
        for (uint32_t k = lane; k < len; k += 32) pos[o + k] = base | src[k];
        o += len;
    }
}

// Range-major flat form of the light rows: flat row k = container row rows[k] (NULL: row k); its values at positions
// [q * range_bits, (q + 1) * range_bits) go to pos[off[q * (n + 1) + k] ...].  Lane q of the row's warp owns range q (at
// most 32 ranges): it holds where the range's part of this row starts and how many values it has received.  The 32
// values a warp reads at a time are dealt out range by range (usually one or two distinct ranges: the lists ascend),
// but nothing here relies on their order -- a list block built from unsorted input lands in the right ranges too.
constexpr uint32_t STREAM_MAX_RANGES = 32;
__global__ void __launch_bounds__(256) flatten_ranges_kernel(const SparseView v, const uint32_t* rows, uint32_t n, uint32_t n_ranges,
                                                             uint32_t range_bits, const uint64_t* off, uint32_t* pos) {
    const uint32_t k_row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k_row >= n) return;
    const uint32_t row = rows ? rows[k_row] : k_row;
    uint64_t cursor = lane < n_ranges ? off[(uint64_t)lane * (n + 1) + k_row] : 0;     // next free slot of range `lane`
    for (uint32_t b = v.row_ptr[row]; b < v.row_ptr[row + 1]; ++b) {
        const uint32_t len = v.blk_len[b];                         // (light rows hold no bitmap blocks)
        const uint32_t hi = v.blk_id[b] << 16;
        const uint16_t* src = v.lists + v.blk_off[b];
        for (uint32_t k0 = 0; k0 < len; k0 += 32) {                // (warp-uniform trip count)
            const bool valid = k0 + lane < len;
            const uint32_t p = hi | (valid ? src[k0 + lane] : 0u);
            const uint32_t q = min(p / range_bits, n_ranges - 1u);
            uint32_t todo = __ballot_sync(0xffffffffu, valid);
            while (todo) {                                         // one round per distinct range among the 32 values
                const uint32_t ql = __shfl_sync(0xffffffffu, q, __ffs(todo) - 1);
                const uint32_t m = __ballot_sync(0xffffffffu, valid && q == ql);
                const uint64_t at = __shfl_sync(0xffffffffu, cursor, (int)ql);
                if (valid && q == ql) pos[at + __popc(m & ((1u << lane) - 1u))] = p;
                if (lane == ql) cursor += __popc(m);
                todo &= ~m;
            }
        }
    }
}

template <int G>
__global__ void __launch_bounds__(SP_MAX_THREADS, 1) sparse_flat_kernel(const SparseJob job, const uint32_t bm_words) {
    extern __shared__ __align__(16) uint32_t s_bits[];            // bm_words: row i as a bitmap
    __shared__ unsigned long long warp_part[SP_MAX_WARPS];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t i = job.i0 + job.shard + (uint64_t)blockIdx.x * job.n_shards;
    if (i >= job.i1) return;
    uint64_t jbeg = job.j0;
    if (job.strict_upper && jbeg < i + 1) jbeg = i + 1;
    const uint64_t js0 = jbeg + (uint64_t)blockIdx.y * FLAT_SLICE;
    if (js0 >= job.j1) return;
    const uint64_t js1 = min(js0 + (uint64_t)FLAT_SLICE, job.j1);
    const uint64_t a0 = job.A.pos_off[i], a1 = job.A.pos_off[i + 1];
    if (a0 == a1) return;                                          // empty row: all counts 0 (out is pre-zeroed)

    uint4* z = reinterpret_cast<uint4*>(s_bits);
    for (uint32_t k = tid; k < bm_words / 4; k += blockDim.x) z[k] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (uint64_t k = a0 + tid; k < a1; k += blockDim.x) {
        const uint32_t v = job.A.pos[k];
        atomicOr(&s_bits[v >> 5], 1u << (v & 31));
    }
    __syncthreads();

    constexpr uint32_t SUBS = 32 / G;                             // partner rows per warp step
    const uint32_t l = lane % G, n_sub = (blockDim.x >> 5) * SUBS;
    const uint64_t* __restrict__ off = job.B.pos_off;
    const uint32_t* __restrict__ pos = job.B.pos;
    unsigned long long acc = 0;
    for (uint64_t jw = js0 + (uint64_t)warp * SUBS; jw < js1; jw += n_sub) {      // warp-uniform trip count
        const uint64_t j = jw + lane / G;
        const bool valid = j < js1;
        uint32_t c = 0;
        if (valid) {
            const uint64_t o0 = off[j], o1 = off[j + 1];
            for (uint64_t k = o0 + l; k < o1; k += G) {
                const uint32_t v = __ldg(pos + k);
                c += (s_bits[v >> 5] >> (v & 31)) & 1u;
            }
        }
        acc += c;
        if (job.out) {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (valid && l == 0) job.out[(i - job.i0) * job.ld + (j - job.j0)] = c;
        }
    }
    if (job.total) {
        const unsigned long long w = warp_sum(acc);
        if (lane == 0) warp_part[warp] = w;
        __syncthreads();
        if (tid == 0) {
            unsigned long long t = 0;
            for (uint32_t k = 0; k < (blockDim.x >> 5); ++k) t += warp_part[k];
            if (t) atomicAdd(job.total, t);
        }
    }
}

// ---- row-group stream kernel: totals over rows that hold few values ---------------------------------------
// For a TOTAL the identity of the partner row does not matter, only how many rows i of a group contain each of its
// positions.  A CTA takes a group of up to 32 consecutive rows i (at most STREAM_ENTRIES values together; groups
// are cut on the host, StormState::h_group_start), builds position -> row-mask in a shared-memory hash table
// (open addressing, load <= 1/2), and then streams the positions of ALL later rows -- one contiguous range of the
// flat mirror, whatever rows they belong to -- through it: total += popc(mask[pos]).  Every (i, j, common position)
// triple is still counted individually (it is the list -> bitmap probe of storm.c:632-646 against 32 rows at
// once); per partner position there is one coalesced 4-byte load and one or two shared-memory probes for up to
// 32 pairs.  Pairs inside a group (j in the group, i < j) take the same table with the mask cut to the rows below j.
constexpr uint32_t STREAM_CAP = 16384;           // hash slots (keys + masks: 128 KiB)
constexpr uint32_t STREAM_ENTRIES = STREAM_CAP / 2;
constexpr uint32_t STREAM_GROUP = 32;            // rows per group (one mask bit each)
constexpr uint64_t STREAM_SLICE = 1u << 18;      // partner positions per CTA

struct StreamJob {
    const uint64_t* a_off; const uint32_t* a_pos;     // flat form of the rows i (CSR offsets, absolute positions)
    const uint64_t* b_off; const uint32_t* b_pos;     // flat form of the partner rows j
    const uint32_t* group_start;   // device: first row of each group of A rows, n_groups + 1 entries
    uint32_t g0, n_groups_job;     // groups [g0, g0 + n_groups_job) overlap rows [i0, i1)
    uint64_t i0, i1, j0, j1;
    int strict_upper;
    uint32_t shard, n_shards;      // groups are dealt round-robin to shards
    unsigned long long* total;
};

static_assert(STREAM_CAP == (1u << 14) && STREAM_GROUP <= 32, "stream_slot keeps 14 bits; a row mask has 32");
__device__ __forceinline__ uint32_t stream_slot(uint32_t p) { return (p * 2654435761u) >> 18; }   // 14 bits
// In front of the table: a hashed bit filter of 2^18 bits (32 KiB).  Nearly every partner position is in none of
// the group's rows; the filter answers that with one shared-memory load and no loop (at most 8192 of its bits are
// set: 3 % false positives, which simply go on to the table).
constexpr uint32_t STREAM_FILTER_WORDS = (1u << 18) / 32;
__device__ __forceinline__ uint32_t stream_filter_bit(uint32_t p) { return (p * 2654435761u) >> 14; }   // 18 bits

__device__ __forceinline__ uint32_t stream_lookup(const uint32_t* keys, const uint32_t* masks, uint32_t p) {
    const uint32_t f = stream_filter_bit(p);
    if (!((masks[STREAM_CAP + (f >> 5)] >> (f & 31)) & 1u)) return 0u;         // filter words follow the masks
    uint32_t slot = f >> 4;                                                    // = stream_slot(p)
    for (;;) {
        const uint32_t k = keys[slot];
        if (k == p + 1u) return masks[slot];
        if (k == 0u) return 0u;
        slot = (slot + 1u) & (STREAM_CAP - 1u);
    }
}

__global__ void __launch_bounds__(SP_MAX_THREADS, 1) sparse_stream_kernel(const StreamJob job) {
    extern __shared__ __align__(16) uint32_t s_tab[];             // keys[STREAM_CAP] | masks[STREAM_CAP] | filter[STREAM_FILTER_WORDS]
    __shared__ uint64_t s_off[STREAM_GROUP + 1];
    __shared__ unsigned long long warp_part[SP_MAX_WARPS];
    uint32_t* keys = s_tab;
    uint32_t* masks = s_tab + STREAM_CAP;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const uint32_t gq = job.shard + blockIdx.x * job.n_shards;
    if (gq >= job.n_groups_job) return;
    const uint32_t g = job.g0 + gq;
    const uint64_t ib = max((uint64_t)job.group_start[g], job.i0), ie = min((uint64_t)job.group_start[g + 1], job.i1);
    if (ib >= ie) return;
    const uint32_t R = (uint32_t)(ie - ib);
    // partner rows all of whose pairs with the group count: [js, j1); rows inside the group: [jin0, jin1)
    uint64_t js = job.j0, jin0 = 0, jin1 = 0;
    if (job.strict_upper) {
        js = max(job.j0, ie);
        jin0 = max(job.j0, ib + 1);
        jin1 = min(job.j1, ie);
    }
    const uint64_t p_beg = js < job.j1 ? job.b_off[js] : 0, p_end = js < job.j1 ? job.b_off[job.j1] : 0;
    const uint64_t k0 = p_beg + (uint64_t)blockIdx.y * STREAM_SLICE;
    const uint64_t k1 = min(k0 + STREAM_SLICE, p_end);
    const bool inner = blockIdx.y == 0 && jin0 < jin1;
    if (k0 >= k1 && !inner) return;

    uint4* z = reinterpret_cast<uint4*>(s_tab);
    for (uint32_t k = tid; k < (2 * STREAM_CAP + STREAM_FILTER_WORDS) / 4; k += blockDim.x) z[k] = make_uint4(0, 0, 0, 0);
    if (tid <= R) s_off[tid] = job.a_off[ib + tid];
    __syncthreads();
    const uint64_t e0 = s_off[0], e1 = s_off[R];
    for (uint64_t e = e0 + tid; e < e1; e += blockDim.x) {
        uint32_t r = 0;                                            // row of entry e: largest r with s_off[r] <= e
#pragma unroll
        for (uint32_t step = 16; step > 0; step >>= 1)
            if (r + step < R && s_off[r + step] <= e) r += step;
        const uint32_t p = job.a_pos[e];
        const uint32_t f = stream_filter_bit(p);
        atomicOr(&masks[STREAM_CAP + (f >> 5)], 1u << (f & 31));
        uint32_t slot = stream_slot(p);
        for (;;) {
            const uint32_t old = atomicCAS(&keys[slot], 0u, p + 1u);
            if (old == 0u || old == p + 1u) { atomicOr(&masks[slot], 1u << r); break; }
            slot = (slot + 1u) & (STREAM_CAP - 1u);
        }
    }
    __syncthreads();

    unsigned long long acc = 0;
    const uint32_t* __restrict__ pos = job.b_pos;
    {   // rows beyond the group: every row of the group pairs with them.  16-byte loads, eight positions in flight
        // per thread: the loop is bound by the latency of the dependent shared-memory probes.
        auto hits = [&](uint32_t p) { return (unsigned)__popc(stream_lookup(keys, masks, p)); };
        const uint64_t ka = min(k1, (uint64_t)((k0 + 3ull) & ~3ull));           // first 16-byte aligned position
        for (uint64_t k = k0 + tid; k < ka; k += blockDim.x) acc += hits(__ldg(pos + k));
        const uint4* __restrict__ p4 = reinterpret_cast<const uint4*>(pos + ka);
        const uint64_t n4 = (k1 - ka) >> 2;
        uint64_t q = tid;
        for (; q + blockDim.x < n4; q += 2ull * blockDim.x) {
            const uint4 a = __ldg(p4 + q), b = __ldg(p4 + q + blockDim.x);
            acc += hits(a.x) + hits(a.y) + hits(a.z) + hits(a.w) + hits(b.x) + hits(b.y) + hits(b.z) + hits(b.w);
        }
        for (; q < n4; q += blockDim.x) {
            const uint4 a = __ldg(p4 + q);
            acc += hits(a.x) + hits(a.y) + hits(a.z) + hits(a.w);
        }
        for (uint64_t k = ka + 4ull * n4 + tid; k < k1; k += blockDim.x) acc += hits(__ldg(pos + k));
    }
    if (inner) {   // rows of the group itself (same container): row j pairs with the group's rows below it
        for (uint64_t j = jin0 + warp; j < jin1; j += (blockDim.x >> 5)) {
            const uint32_t below = (1u << (uint32_t)(j - ib)) - 1u;                 // j - ib in [1, 31]
            for (uint64_t k = job.b_off[j] + lane; k < job.b_off[j + 1]; k += 32)
                acc += __popc(stream_lookup(keys, masks, __ldg(pos + k)) & below);
        }
    }
    const unsigned long long w = warp_sum(acc);
    if (lane == 0) warp_part[warp] = w;
    __syncthreads();
    if (tid == 0) {
        unsigned long long t = 0;
        for (uint32_t k = 0; k < (blockDim.x >> 5); ++k) t += warp_part[k];
        if (t) atomicAdd(job.total, t);
    }
}

// Dense route: one CTA per row writes the row's blocks into a zeroed row-major arena in the
// layout of the contiguous model (bit v -> word v/64, bit v%64; storm.c:1114), after which
// the query is the dense tile kernel's.  A bitmap block is a 8 KiB copy; a list block sets
// its bits with 64-bit atomic ORs (two values of a list may share a word).
__global__ void __launch_bounds__(256) densify_rows_kernel(const SparseView v, uint64_t* dense, uint64_t stride, uint32_t row0) {
    const uint32_t row = row0 + blockIdx.x;                      // dense row blockIdx.x = container row row0 + blockIdx.x
    unsigned long long* out = reinterpret_cast<unsigned long long*>(dense + (uint64_t)blockIdx.x * stride);
    for (uint32_t b = v.row_ptr[row]; b < v.row_ptr[row + 1]; ++b) {
        const uint32_t len = v.blk_len[b];
        unsigned long long* blk = out + (uint64_t)v.blk_id[b] * BLOCK_WORDS;
        if (len & BITMAP_FLAG) {
            const uint64_t* src = v.words + v.blk_off[b];
            for (uint32_t w = threadIdx.x; w < BLOCK_WORDS; w += blockDim.x) blk[w] = src[w];
        } else {
            const uint16_t* src = v.lists + v.blk_off[b];
            for (uint32_t k = threadIdx.x; k < len; k += blockDim.x) {
                const uint32_t x = src[k];
                atomicOr(blk + (x >> 6), 1ull << (x & 63));
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// device mirror
// ---------------------------------------------------------------------------------
struct StormState {
    int device = -1;
    cudaStream_t stream = nullptr;
    bool dirty = true;
    uint32_t n_rows = 0, max_blocks = 0;
    uint32_t max_blk_id = 0;             // largest block index of any row (row width of the dense form)
    uint64_t total_nnz = 0, total_blocks = 0;
    uint64_t* d_dense = nullptr; uint64_t dense_cap_words = 0; bool dense_valid = false;   // densified rows (dense route)
    int last_route = 0;                  // 1 = sparse kernel, 2 = densified + dense tile kernel
    uint32_t *d_row_ptr = nullptr, *d_row_nnz = nullptr, *d_blk_id = nullptr, *d_blk_len = nullptr;
    uint64_t* d_blk_off = nullptr; uint16_t* d_lists = nullptr; uint64_t* d_words = nullptr;
    unsigned long long* d_total = nullptr; unsigned long long* h_total = nullptr;
    uint64_t n_bitmap_blocks = 0;        // blocks held as bitmaps (none: the flat probe kernel applies)
    uint64_t* d_pos_off = nullptr; uint32_t* d_pos = nullptr; bool flat_valid = false;     // flat form (ensure_flat)
    uint32_t max_row_nnz = 0;
    std::vector<uint32_t> h_group_start; // row groups of the stream kernel (<= 32 rows, <= STREAM_ENTRIES values)
    uint32_t* d_group_start = nullptr;
    // Split route (containers that hold heavy rows -- a bitmap block, or more values than a row group may -- among
    // light ones): the light rows' own flat form and row groups, and the list of heavy rows.
    uint32_t n_light = 0, n_heavy = 0; uint64_t light_nnz = 0;
    uint32_t n_ranges = 1, range_bits = 0;       // position ranges of the light mirror (range q = positions [q, q + 1) * range_bits)
    bool light_mirror = false;                   // the arrays below exist (heavy rows, or more than one range)
    std::vector<uint64_t> range_nnz;
    uint32_t *d_light_rows = nullptr, *d_heavy_rows = nullptr, *d_lgroup_start = nullptr, *d_lpos = nullptr;
    uint64_t* d_lpos_off = nullptr; bool lflat_valid = false;
    std::vector<uint32_t> h_lgroup_start;
    // Whole-container queries run on the device set (devices.h): this state is the replica on the set's first device
    // and owns the replicas on the others (resolved at the first query; rectangles and XY^T stay on this one).
    std::vector<StormState*> replicas;
    bool set_resolved = false;
    uint64_t* d_band[2] = {nullptr, nullptr}; uint64_t band_cap_words = 0;   // arenas of the banded dense route
};

inline StormState* state_of(const STORM_t* s) { return static_cast<StormState*>(s->b200); }

void free_mirror(StormState* st) {
    for (void* p : {(void*)st->d_row_ptr, (void*)st->d_row_nnz, (void*)st->d_blk_id, (void*)st->d_blk_len,
                    (void*)st->d_blk_off, (void*)st->d_lists, (void*)st->d_words, (void*)st->d_pos_off, (void*)st->d_pos,
                    (void*)st->d_group_start, (void*)st->d_light_rows, (void*)st->d_heavy_rows, (void*)st->d_lgroup_start,
                    (void*)st->d_lpos, (void*)st->d_lpos_off})
        if (p) cudaFree(p);
    st->d_light_rows = st->d_heavy_rows = st->d_lgroup_start = st->d_lpos = nullptr; st->d_lpos_off = nullptr;
    st->lflat_valid = false; st->h_lgroup_start.clear(); st->light_mirror = false;
    st->d_row_ptr = st->d_row_nnz = st->d_blk_id = st->d_blk_len = nullptr;
    st->d_blk_off = nullptr; st->d_lists = nullptr; st->d_words = nullptr;
    st->d_pos_off = nullptr; st->d_pos = nullptr; st->flat_valid = false;
    st->d_group_start = nullptr; st->h_group_start.clear();
    st->dense_valid = false;
}

template <typename T>
int upload(T** dst, const std::vector<T>& src, cudaStream_t stream) {
    const size_t n = std::max<size_t>(src.size(), 1);
    STORM_CUDA_TRY(cudaMalloc(dst, n * sizeof(T)));
    if (!src.empty()) STORM_CUDA_TRY(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
    return STORM_B200_OK;
}

int ensure_state(StormState* st) {
    if (st->device >= 0 && st->d_total) return STORM_B200_OK;
    int rc = require_device();
    if (rc) return rc;
    STORM_CUDA_TRY(cudaGetDevice(&st->device));                     // (callers made the state's device current, if it has one)
    STORM_CUDA_TRY(cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking));
    STORM_CUDA_TRY(cudaMalloc(&st->d_total, 8));
    STORM_CUDA_TRY(cudaMallocHost(&st->h_total, 8));
    return STORM_B200_OK;
}

// The host containers flattened (CSR of blocks + one u16 pool + one bitmap pool): built once per change of the
// container and uploaded to every replica.
struct HostMirror {
    std::vector<uint32_t> row_ptr, row_nnz, row_nnz_dev, blk_id, blk_len, group_start;
    std::vector<uint32_t> light_rows, heavy_rows, lgroup_start;       // split route (filled only if there are heavy rows)
    std::vector<uint64_t> blk_off, words, pos_off, lpos_off, range_nnz;
    uint64_t light_nnz = 0;
    uint32_t n_ranges = 1, range_bits = 0;
    std::vector<uint16_t> lists;
    uint32_t max_blocks = 0, max_blk_id = 0, max_row_nnz = 0;
    uint64_t n_bitmap_blocks = 0, total_nnz = 0;
};

void build_host_mirror(const STORM_t* s, HostMirror* h) {
    h->row_ptr.assign(s->n_conts + 1, 0);
    h->row_nnz.assign(s->n_conts, 0);
    for (uint32_t r = 0; r < s->n_conts; ++r) {
        const STORM_bitmap_cont_t* row = &s->conts[r];
        h->row_ptr[r] = (uint32_t)h->blk_id.size();
        h->max_blocks = std::max(h->max_blocks, row->n_bitmaps);
        for (uint32_t b = 0; b < row->n_bitmaps; ++b) {
            const STORM_bitmap_t* k = &row->bitmaps[b];
            h->blk_id.push_back(k->id);
            h->max_blk_id = std::max(h->max_blk_id, k->id);
            if (k->n_bitmap) {
                ++h->n_bitmap_blocks;
                h->blk_len.push_back(k->n_bits_set | BITMAP_FLAG);
                h->blk_off.push_back(h->words.size());
                h->words.insert(h->words.end(), k->data, k->data + BLOCK_WORDS);
                h->row_nnz[r] += k->n_bits_set;
            } else {
                while (h->lists.size() % 8) h->lists.push_back(0);       // 16-byte aligned lists
                h->blk_off.push_back(h->lists.size());
                // set semantics on the device: adjacent duplicates (which the reference's list
                // builder keeps, storm.c:548-556) collapse, as they do in a bitmap block
                uint32_t kept = 0;
                for (uint32_t v = 0; v < k->n_scalar; ++v)
                    if (v == 0 || k->scalar[v] != k->scalar[v - 1]) { h->lists.push_back(k->scalar[v]); ++kept; }
                h->blk_len.push_back(kept);
                h->row_nnz[r] += kept;
            }
        }
    }
    h->row_ptr[s->n_conts] = (uint32_t)h->blk_id.size();
    while (h->lists.size() % 8) h->lists.push_back(0);             // 16-byte loads of the last list stay inside the pool
    h->pos_off.assign(s->n_conts + 1, 0);                          // CSR offsets of the flat form (built on demand)
    for (uint32_t r = 0; r < s->n_conts; ++r) h->pos_off[r + 1] = h->pos_off[r] + h->row_nnz[r];
    for (uint32_t v : h->row_nnz) { h->max_row_nnz = std::max(h->max_row_nnz, v); h->total_nnz += v; }
    stream_groups(h->row_nnz.data(), s->n_conts, &h->group_start);
    h->row_nnz_dev = h->row_nnz;                                   // device copy: + the "holds a bitmap block" flag
    for (uint32_t r = 0; r < s->n_conts; ++r)
        for (uint32_t b = h->row_ptr[r]; b < h->row_ptr[r + 1]; ++b)
            if (h->blk_len[b] & BITMAP_FLAG) { h->row_nnz_dev[r] |= ROW_HAS_BITMAP; break; }
    // heavy rows: a bitmap block, or too many values for a row group of the stream kernel
    for (uint32_t r = 0; r < s->n_conts; ++r)
        if ((h->row_nnz_dev[r] & ROW_HAS_BITMAP) || h->row_nnz[r] > STREAM_ENTRIES) { h->row_nnz_dev[r] |= ROW_HEAVY; h->heavy_rows.push_back(r); }
    // The light rows' mirror for the stream kernel, cut into n_ranges position ranges so that 32 rows fit the
    // shared-memory table range by range, at a low load, whatever a row holds in total (stream_ranges).  Here only its
    // shape is fixed (the route model needs it); the arrays are built by build_light_arrays when a route reads them.
    const uint32_t span = h->max_blk_id + 1;
    const uint64_t n_light = s->n_conts - h->heavy_rows.size();
    uint64_t light_total = 0;
    for (uint32_t r = 0; r < s->n_conts; ++r) if (!(h->row_nnz_dev[r] & ROW_HEAVY)) light_total += h->row_nnz[r];
    h->light_nnz = light_total;
    h->n_ranges = n_light ? stream_ranges((double)light_total / (double)n_light, span, &h->range_bits) : 1;
}

// The light mirror's arrays (range-major offsets, row groups, the list of light rows): built only when the route a
// query takes reads them -- a container the cost model densifies never pays for them.
void build_light_arrays(const STORM_t* s, HostMirror* h) {
    const uint64_t n_light = s->n_conts - h->heavy_rows.size();
    {
        const uint32_t P = h->n_ranges;
        if (!h->heavy_rows.empty())
            for (uint32_t r = 0; r < s->n_conts; ++r) if (!(h->row_nnz_dev[r] & ROW_HEAVY)) h->light_rows.push_back(r);
        std::vector<uint32_t> cnt((size_t)n_light * P, 0);           // values of light row k in range q: cnt[k * P + q]
        uint64_t k = 0;
        for (uint32_t r = 0; r < s->n_conts; ++r) {
            if (h->row_nnz_dev[r] & ROW_HEAVY) continue;
            for (uint32_t b = h->row_ptr[r]; b < h->row_ptr[r + 1]; ++b) {
                const uint64_t lo = (uint64_t)h->blk_id[b] << 16;
                const uint16_t* list = h->lists.data() + h->blk_off[b];
                const uint32_t len = h->blk_len[b];
                const uint64_t q0 = std::min<uint64_t>(lo / h->range_bits, P - 1), q1 = std::min<uint64_t>((lo + 65535) / h->range_bits, P - 1);
                if (q0 == q1) { cnt[k * P + q0] += len; continue; }
                for (uint32_t e = 0; e < len; ++e) ++cnt[k * P + std::min<uint64_t>((lo + list[e]) / h->range_bits, P - 1)];
            }
            ++k;
        }
        // range-major CSR: range q's offsets are lpos_off[q * (n_light + 1) ...], absolute into one position array
        h->lpos_off.assign((size_t)P * (n_light + 1), 0);
        h->range_nnz.assign(P, 0);
        uint64_t at = 0;
        for (uint32_t q = 0; q < P; ++q) {
            for (uint64_t i = 0; i < n_light; ++i) { h->lpos_off[q * (n_light + 1) + i] = at; at += cnt[i * P + q]; h->range_nnz[q] += cnt[i * P + q]; }
            h->lpos_off[q * (n_light + 1) + n_light] = at;
        }
        // row groups shared by every range: at most 32 rows, at most STREAM_ENTRIES values in any one range
        std::vector<uint32_t>& gs = h->lgroup_start;
        gs.assign(1, 0u);
        std::vector<uint64_t> in_group(P, 0);
        for (uint64_t i = 0; i < n_light; ++i) {
            bool cut = i > gs.back() && i - gs.back() == STREAM_GROUP;
            for (uint32_t q = 0; q < P && !cut; ++q) cut = i > gs.back() && in_group[q] + cnt[i * P + q] > STREAM_ENTRIES;
            if (cut) { gs.push_back((uint32_t)i); std::fill(in_group.begin(), in_group.end(), 0); }
            for (uint32_t q = 0; q < P; ++q) in_group[q] += cnt[i * P + q];
        }
        gs.push_back((uint32_t)n_light);
    }
}

// One replica's copy of the light mirror's arrays (its positions are flattened on the device on first use).
int upload_light(StormState* st, const HostMirror& h) {
    int rc = STORM_B200_OK;
    st->h_lgroup_start = h.lgroup_start;
    st->range_nnz = h.range_nnz;
    if ((!h.light_rows.empty() && (rc = upload(&st->d_light_rows, h.light_rows, st->stream))) ||
        (rc = upload(&st->d_lgroup_start, h.lgroup_start, st->stream)) || (rc = upload(&st->d_lpos_off, h.lpos_off, st->stream)))
        return rc;
    st->light_mirror = true;
    return STORM_B200_OK;
}

// One replica's copy of the mirror (on the current device = the replica's).
int upload_mirror(StormState* st, const HostMirror& h, uint32_t n_conts) {
    int rc = ensure_state(st);
    if (rc) return rc;
    free_mirror(st);
    st->h_group_start = h.group_start;
    if ((rc = upload(&st->d_pos_off, h.pos_off, st->stream)) || (rc = upload(&st->d_group_start, h.group_start, st->stream)) ||
        (rc = upload(&st->d_row_ptr, h.row_ptr, st->stream)) || (rc = upload(&st->d_row_nnz, h.row_nnz_dev, st->stream)) ||
        (rc = upload(&st->d_blk_id, h.blk_id, st->stream)) || (rc = upload(&st->d_blk_len, h.blk_len, st->stream)) ||
        (rc = upload(&st->d_blk_off, h.blk_off, st->stream)) || (rc = upload(&st->d_lists, h.lists, st->stream)) ||
        (rc = upload(&st->d_words, h.words, st->stream)))
        return rc;
    st->n_heavy = (uint32_t)h.heavy_rows.size();
    st->n_light = n_conts - st->n_heavy;
    st->light_nnz = h.light_nnz;
    st->n_ranges = h.n_ranges; st->range_bits = h.range_bits;
    if (st->n_heavy && (rc = upload(&st->d_heavy_rows, h.heavy_rows, st->stream))) return rc;
    if (!h.lpos_off.empty() && (rc = upload_light(st, h))) return rc;
    STORM_CUDA_TRY(cudaStreamSynchronize(st->stream));              // (the host vectors may die after this)
    st->n_rows = n_conts;
    st->max_blocks = h.max_blocks;
    st->max_blk_id = h.max_blk_id;
    st->max_row_nnz = h.max_row_nnz;
    st->total_nnz = h.total_nnz;
    st->total_blocks = h.blk_id.size();
    st->n_bitmap_blocks = h.n_bitmap_blocks;
    st->dirty = false;
    return STORM_B200_OK;
}

// The replicas of a container: the device set in force at its first query, first device = this state's.
int resolve_replicas(StormState* st) {
    if (st->set_resolved) return STORM_B200_OK;
    std::vector<int> ids;
    int rc = query_devices(&ids);
    if (rc) return rc;
    if (!st->d_total) st->device = ids[0];                           // (not initialised yet: adopt the set's first device)
    size_t own = 0;                                                  // the entry of the set this state stands for
    for (size_t k = 0; k < ids.size(); ++k) if (ids[k] == st->device) { own = k; break; }
    for (size_t k = 0; k < ids.size(); ++k) {
        if (k == own) continue;
        StormState* rep = new (std::nothrow) StormState();
        if (!rep) { set_error("out of host memory"); return STORM_B200_ENOMEM; }
        rep->device = ids[k];
        rep->set_resolved = true;
        st->replicas.push_back(rep);
    }
    st->set_resolved = true;
    return STORM_B200_OK;
}

// Flatten the host containers and bring every replica up to date (whole-container rebuild on change).
// `all` = false: this state only (rectangles, XY^T).
int choose_route(const StormState* st, uint64_t n_rows);
bool route_reads_light(int route, const StormState* st);

int sync_mirror(const STORM_t* s, StormState* st, bool all = false) {
    if (all) { int rc = resolve_replicas(st); if (rc) return rc; }
    std::vector<StormState*> targets{st};
    if (all) targets.insert(targets.end(), st->replicas.begin(), st->replicas.end());
    bool need = false;
    for (StormState* t : targets) need = need || t->dirty || !t->d_total || t->n_rows != s->n_conts;
    if (!need) return STORM_B200_OK;
    HostMirror h;
    build_host_mirror(s, &h);
    if (all && s->n_conts >= 2) {                                        // whole-container queries: will the route read the light mirror?
        StormState probe;                                                // (the statistics choose_route looks at, nothing else)
        probe.n_heavy = (uint32_t)h.heavy_rows.size(); probe.n_light = s->n_conts - probe.n_heavy; probe.light_nnz = h.light_nnz;
        probe.n_ranges = h.n_ranges; probe.max_blk_id = h.max_blk_id; probe.max_blocks = h.max_blocks; probe.max_row_nnz = h.max_row_nnz;
        probe.total_nnz = h.total_nnz; probe.total_blocks = h.blk_id.size(); probe.n_bitmap_blocks = h.n_bitmap_blocks;
        if (route_reads_light(choose_route(&probe, s->n_conts), &probe)) build_light_arrays(s, &h);
    }
    for (StormState* t : targets) {
        if (!(t->dirty || !t->d_total || t->n_rows != s->n_conts)) continue;
        DeviceGuard guard(t->device);
        int rc = upload_mirror(t, h, s->n_conts);
        if (rc) return rc;
    }
    return STORM_B200_OK;
}

// The light mirror's arrays after the fact (a route knob changed since the mirror was built): every replica that lacks them.
int sync_light(const STORM_t* s, StormState* st) {
    std::vector<StormState*> targets{st};
    targets.insert(targets.end(), st->replicas.begin(), st->replicas.end());
    bool need = false;
    for (StormState* t : targets) need = need || !t->light_mirror;
    if (!need) return STORM_B200_OK;
    HostMirror h;
    build_host_mirror(s, &h);
    build_light_arrays(s, &h);
    for (StormState* t : targets) {
        if (t->light_mirror) continue;
        DeviceGuard guard(t->device);
        int rc = upload_light(t, h);
        if (rc) return rc;
        STORM_CUDA_TRY(cudaStreamSynchronize(t->stream));               // (the host vectors die at scope exit)
    }
    return STORM_B200_OK;
}

SparseView view_of(const StormState* st) {
    return SparseView{st->d_row_ptr, st->d_row_nnz, st->d_blk_id, st->d_blk_len, st->d_blk_off, st->d_lists, st->d_words, st->n_rows,
                      st->flat_valid ? st->d_pos_off : nullptr, st->flat_valid ? st->d_pos : nullptr, st->max_blk_id + 1,
                      st->n_rows ? (float)((double)st->total_nnz / (double)st->n_rows) : 0.0f};
}

int g_sparse_flat = 1;   // STORM_b200_set_sparse_flat(0): always the block kernel

// The flat probe kernel applies when no block is a bitmap and a whole row fits shared memory as a bitmap.
bool flat_eligible(const StormState* st) {
    return g_sparse_flat && st->n_bitmap_blocks == 0 && ((uint64_t)st->max_blk_id + 1) * 8192 <= FLAT_MAX_SMEM;
}

// Build the flat form (absolute positions, CSR) of an eligible container on the device; a no-op otherwise.
int ensure_flat(StormState* st) {
    if (st->flat_valid || !flat_eligible(st) || st->n_rows == 0) return STORM_B200_OK;
    if (!st->d_pos) {
        if (cudaMalloc(&st->d_pos, std::max<uint64_t>(st->total_nnz, 1) * sizeof(uint32_t)) != cudaSuccess) {
            cudaGetLastError();
            st->d_pos = nullptr;
            return STORM_B200_OK;                                  // no room: the block kernel answers
        }
    }
    flatten_rows_kernel<<<(st->n_rows + 7) / 8, 256, 0, st->stream>>>(view_of(st), nullptr, st->n_rows, st->d_pos_off, st->d_pos);
    STORM_CUDA_TRY(cudaGetLastError());
    count_launch();
    st->flat_valid = true;
    return STORM_B200_OK;
}

// The light rows' flat form (split route).  Returns false if there is no room for it.
int ensure_light_flat(StormState* st, bool* ok) {
    *ok = true;
    if (st->lflat_valid) return STORM_B200_OK;
    if (!st->d_lpos && cudaMalloc(&st->d_lpos, std::max<uint64_t>(st->light_nnz, 1) * sizeof(uint32_t)) != cudaSuccess) {
        cudaGetLastError();
        st->d_lpos = nullptr; *ok = false;
        return STORM_B200_OK;
    }
    flatten_ranges_kernel<<<(st->n_light + 7) / 8, 256, 0, st->stream>>>(view_of(st), st->d_light_rows, st->n_light, st->n_ranges, st->range_bits, st->d_lpos_off, st->d_lpos);
    STORM_CUDA_TRY(cudaGetLastError());
    count_launch();
    st->lflat_valid = true;
    return STORM_B200_OK;
}

template <int G>
int launch_flat_g(const SparseJob& job, uint32_t bm_words, dim3 grid, cudaStream_t stream) {
    const size_t smem = (size_t)bm_words * 4;
    STORM_CUDA_TRY(cudaFuncSetAttribute(sparse_flat_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLAT_MAX_SMEM));
    sparse_flat_kernel<G><<<grid, smem > SP_ONE_CTA_SMEM / 2 ? SP_MAX_THREADS : SP_MAX_THREADS / 2, smem, stream>>>(job, bm_words);
    STORM_CUDA_TRY(cudaGetLastError());
    count_launch();
    return STORM_B200_OK;
}

int launch_flat(const SparseJob& job, cudaStream_t stream) {
    const uint32_t span = std::max(job.A.n_blk_span, job.B.n_blk_span);       // every probed position is below span * 65536
    const uint32_t bm_words = span * 2048;
    const uint64_t rows_i = (job.i1 - job.i0 + job.n_shards - 1 - job.shard) / job.n_shards;
    if (rows_i == 0) return STORM_B200_OK;
    const uint64_t slices = (job.j1 - job.j0 + FLAT_SLICE - 1) / FLAT_SLICE;
    if (slices > 65535) { set_error("too many partner rows for one launch (%llu)", (unsigned long long)(job.j1 - job.j0)); return STORM_B200_EINVAL; }
    dim3 grid((unsigned)rows_i, (unsigned)slices);
    // lanes per partner row: about four values per lane
    const float a = job.B.avg_nnz;
    if (a <= 4.f) return launch_flat_g<1>(job, bm_words, grid, stream);
    if (a <= 8.f) return launch_flat_g<2>(job, bm_words, grid, stream);
    if (a <= 16.f) return launch_flat_g<4>(job, bm_words, grid, stream);
    if (a <= 32.f) return launch_flat_g<8>(job, bm_words, grid, stream);
    if (a <= 64.f) return launch_flat_g<16>(job, bm_words, grid, stream);
    return launch_flat_g<32>(job, bm_words, grid, stream);
}

int g_sparse_stream = 1;   // STORM_b200_set_sparse_flat(2): totals of light rows through the row-group stream kernel

// Totals of rows [i0, i1) of `a` against rows [j0, j1) of `b` with the row-group stream kernel.  Needs both flat
// forms and no row of `a` above STREAM_ENTRIES values.
bool stream_eligible(const StormState* a, const StormState* b) {
    return g_sparse_flat && g_sparse_stream && a->flat_valid && b->flat_valid && a->max_row_nnz <= STREAM_ENTRIES &&
           a->d_group_start != nullptr;
}

}  // namespace

// Row groups of the stream kernel: consecutive rows, at most 32 and at most STREAM_ENTRIES values together.
// Returns false if a single row holds more than a group may.
bool stream_groups(const uint32_t* row_nnz, uint64_t n_rows, std::vector<uint32_t>* group_start) {
    std::vector<uint32_t>& gs = *group_start;
    gs.assign(1, 0u);
    uint64_t in_group = 0;
    bool ok = true;
    for (uint64_t r = 0; r < n_rows; ++r) {
        if (row_nnz[r] > STREAM_ENTRIES) ok = false;
        if (r > gs.back() && (r - gs.back() == STREAM_GROUP || in_group + row_nnz[r] > STREAM_ENTRIES)) { gs.push_back((uint32_t)r); in_group = 0; }
        in_group += row_nnz[r];
    }
    gs.push_back((uint32_t)n_rows);
    return ok;
}

// Position ranges the light mirror is cut into: a group should hold 32 rows at a table load near 1/4, i.e. a row about
// 128 values per range; at most 32 ranges (one lane each in flatten_ranges_kernel).  Ranges are equal spans of
// `range_bits` positions and need not respect block boundaries.
uint32_t stream_ranges(double avg_nnz, uint32_t span, uint32_t* range_bits) {
    const uint64_t bits = (uint64_t)std::max(1u, span) << 16;
    const double need = avg_nnz / 128.0;
    static_assert(STREAM_MAX_RANGES <= 32, "flatten_ranges_kernel keeps one range per lane");
    const uint32_t p = need <= 1.0 ? 1u : (uint32_t)std::min((double)STREAM_MAX_RANGES, std::ceil(need));
    const uint64_t rb = (bits + p - 1) / p;
    if (range_bits) *range_bits = (uint32_t)std::min<uint64_t>(rb, 0xFFFFFFFFull);
    return (uint32_t)((bits + rb - 1) / rb);
}

// Seconds the stream kernel needs for `pairs` pairs of rows holding avg_nnz values: one probe per partner position
// and GROUP of rows i (a group holds min(32, 8192 / values per row and range) rows); a probe is one filter lookup plus,
// for the hits and the filter's false positives, a walk through the table, both growing with the table's load.  Fitted to
// 10 000 x 524 288 at 5 / 104 / 300 / 524 / 1 000 / 1 500 / 2 097 values per row on whole-block ranges (0.03 / 0.22 / 0.65 /
// 1.35 / 3.58 / 4.12 / 7.63 ms with 1 / 1 / 2 / 3 / 4 / 8 / 8 position ranges, profiles/r02_sparse_routes_ranges.jsonl): 1.06 ps
// per probe on a nearly empty table, 2.3 ps at load 1/2; with ranges of ~128 values per row it predicts 0.61 / 1.07 / 2.16 /
// 3.23 / 4.48 / 6.44 ms for 300 ... 3 000 values per row against 0.62 / 1.20 / 2.27 / 3.18 / 4.56 / 6.60 measured (_v2).  With `ranges`
// position ranges the same probes run against tables that hold 1 / ranges of each row, one launch per range.
double stream_seconds(double pairs, double avg_nnz, uint32_t ranges) {
    const double per_range = std::max(1.0, avg_nnz / (double)std::max(1u, ranges));
    const double group = std::min(32.0, std::max(1.0, std::floor((double)STREAM_ENTRIES / per_range)));
    const double load = std::min(0.5, group * per_range / (double)STREAM_CAP);
    return 1.5e-5 + 4e-6 * (double)(std::max(1u, ranges) - 1u) + pairs * avg_nnz / group * (1.06e-12 + 4.9e-12 * load * load);
}

// Totals of rows [i0, i1) (flat form a_*, row groups h/d_group_start) against rows [j0, j1) (flat form b_*).
int launch_sparse_stream(const uint64_t* a_off, const uint32_t* a_pos, const std::vector<uint32_t>& gs, const uint32_t* d_group_start,
                         const uint64_t* b_off, const uint32_t* b_pos, uint64_t b_total_nnz,
                         uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, int strict_upper,
                         uint32_t shard, uint32_t n_shards, unsigned long long* d_total, cudaStream_t stream) {
    if (i1 <= i0 || j1 <= j0) return STORM_B200_OK;
    const uint32_t g0 = (uint32_t)(std::upper_bound(gs.begin(), gs.end(), (uint32_t)i0) - gs.begin()) - 1;   // groups overlapping [i0, i1)
    const uint32_t g1 = (uint32_t)(std::lower_bound(gs.begin(), gs.end(), (uint32_t)i1) - gs.begin());
    StreamJob job{};
    job.a_off = a_off; job.a_pos = a_pos; job.b_off = b_off; job.b_pos = b_pos;
    job.group_start = d_group_start;
    job.g0 = g0; job.n_groups_job = g1 - g0;
    job.i0 = i0; job.i1 = i1; job.j0 = j0; job.j1 = j1;
    job.strict_upper = strict_upper; job.shard = shard; job.n_shards = n_shards;
    job.total = d_total;
    const uint64_t my_groups = (job.n_groups_job + n_shards - 1 - shard) / n_shards;
    if (my_groups == 0) return STORM_B200_OK;
    uint64_t slices = (b_total_nnz + STREAM_SLICE - 1) / STREAM_SLICE;      // upper bound on any group's partner stream
    if (slices == 0) slices = 1;
    if (slices > 65535) { set_error("partner stream too long for one launch (%llu values)", (unsigned long long)b_total_nnz); return STORM_B200_EINVAL; }
    const size_t smem = (2 * STREAM_CAP + STREAM_FILTER_WORDS) * sizeof(uint32_t);
    STORM_CUDA_TRY(cudaFuncSetAttribute(sparse_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sparse_stream_kernel<<<dim3((unsigned)my_groups, (unsigned)slices), SP_MAX_THREADS, smem, stream>>>(job);
    STORM_CUDA_TRY(cudaGetLastError());
    count_launch();
    return STORM_B200_OK;
}

namespace {

int launch_stream(const StormState* a, const StormState* b, uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, int strict_upper,
                  uint32_t shard, uint32_t n_shards, unsigned long long* d_total, cudaStream_t stream) {
    return launch_sparse_stream(a->d_pos_off, a->d_pos, a->h_group_start, a->d_group_start, b->d_pos_off, b->d_pos, b->total_nnz,
                                i0, i1, j0, j1, strict_upper, shard, n_shards, d_total, stream);
}

int launch_sparse(const SparseJob& job_in, uint32_t max_blocks, cudaStream_t stream) {
    SparseJob job = job_in;
    if (job.i1 <= job.i0 || job.j1 <= job.j0) return STORM_B200_OK;
    if (g_sparse_flat && job.A.pos && job.B.pos && !job.i_list &&
        (uint64_t)std::max(job.A.n_blk_span, job.B.n_blk_span) * 8192 <= FLAT_MAX_SMEM) return launch_flat(job, stream);
    job.maxb = std::max<uint32_t>(1, std::min<uint32_t>(max_blocks, SP_MAXB_CAP));
    const size_t smem = (size_t)job.maxb * 8192;
    STORM_CUDA_TRY(cudaFuncSetAttribute(sparse_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SP_MAXB_CAP * 8192)));
    const uint64_t rows_i = (job.i1 - job.i0 + job.n_shards - 1 - job.shard) / job.n_shards;   // rows of this shard
    if (rows_i == 0) return STORM_B200_OK;
    // few rows i (the heavy rows of the split route): thinner slices, so that the launch still covers the SMs
    job.slice = SP_SLICE;
    while (job.slice > 64 && rows_i * ((job.j1 - job.j0 + job.slice - 1) / job.slice) < 4ull * 148) job.slice /= 2;
    const uint64_t slices = (job.j1 - job.j0 + job.slice - 1) / job.slice;
    if (slices > 65535) { set_error("too many partner rows for one launch (%llu)", (unsigned long long)(job.j1 - job.j0)); return STORM_B200_EINVAL; }
    dim3 grid((unsigned)rows_i, (unsigned)slices);
    sparse_pairs_kernel<<<grid, smem > SP_ONE_CTA_SMEM ? SP_MAX_THREADS : SP_MAX_THREADS / 2, smem, stream>>>(job);
    STORM_CUDA_TRY(cudaGetLastError());
    count_launch();
    return STORM_B200_OK;
}

// Route of a whole-container query: a cost model, not the reference's CPU-tuned 4096 / 200 constants, and it
// cannot change a result.  The dense tile kernel costs W / 6e13 s per pair whatever the density (FP4 tensor
// form; 3.5e13 for the int8 form) plus one pass over the rows to densify them.  The merge-probe kernel costs,
// per pair, a fixed part, a part per block of the row (block-id merge + dispatch) and a part per value probed:
// 0.08 + 0.16 blocks + 0.0003 values ns, fitted to 10 000 x 524 288 at 104 / 5 242 values per row and
// 3 000 x 1 048 576 at 10 486 (profiles/r01_sparse_timing.jsonl).  Containers without bitmap blocks take the
// row-group stream kernel instead (stream_seconds), which wins up to several hundred values per row: at 10 000 x
// 524 288 the crossover with the tensor kernel is near 700 values per row (0.13 % density).
int g_storm_route = 0;   // 0 auto, 1 sparse kernels, 2 densify + dense tile kernel, 3 split (STORM_b200_set_storm_route)

// The model itself (pure arithmetic: tests/test_abi.py pins its decisions on the measured cases without a GPU).
//   dense  = N(N-1)/2 x W / tensor rate (6e13 wp/s FP4 form, 3.5e13 int8 form) + launch + the densify pass
//   sparse = the cheaper of the block merge/probe kernel and, where it applies, the row-group stream kernel
void storm_route_model(uint64_t n_rows, uint64_t W, double avg_nnz, double avg_blocks, bool stream_applies, bool fp4,
                       bool dense_resident, double* dense_s, double* sparse_s) {
    const double pairs = 0.5 * (double)n_rows * (double)(n_rows - 1);
    const double dense_rate = fp4 ? 6.0e13 : 3.5e13;
    *dense_s = pairs * (double)W / dense_rate + 3e-5 + (dense_resident ? 0.0 : (double)n_rows * (double)W * 8.0 / 2e12);
    *sparse_s = pairs * 1e-9 * (0.08 + 0.16 * avg_blocks + 0.0003 * avg_nnz) + 1e-5;             // block merge/probe kernel
    if (stream_applies)                                                                           // row-group stream kernel
        *sparse_s = std::min(*sparse_s, stream_seconds(pairs, avg_nnz, stream_ranges(avg_nnz, (uint32_t)((W + BLOCK_WORDS - 1) / BLOCK_WORDS), nullptr)));
}

// Split route: the light rows among themselves through the stream kernel, every pair with a heavy row through the
// block merge/probe kernel (rows i = the heavy rows).  Heavy x heavy pairs meet in bitmap blocks: 8 KiB per shared
// block from L2, ~2 ns.
double split_seconds(uint64_t n_rows, uint64_t n_heavy, double light_nnz, double total_nnz, double max_blocks, double n_bitmap_blocks,
                     uint32_t span) {
    const double nl = (double)(n_rows - n_heavy), nh = (double)n_heavy;
    const double hh_pairs = 0.5 * nh * (nh - 1.0), h_pairs = nh * nl + hh_pairs;
    double t = 2e-5 + h_pairs * 1e-9 * (0.08 + 0.16 * max_blocks + 0.0003 * total_nnz / (double)n_rows);
    if (n_heavy) t += hh_pairs * (n_bitmap_blocks / nh) * 2e-9;
    if (nl >= 2.0) t += stream_seconds(0.5 * nl * (nl - 1.0), light_nnz / nl, stream_ranges(light_nnz / nl, span, nullptr));
    return t;
}

// Which route a whole-container query takes: 1 sparse kernels, 2 densified rows + tile kernel, 4 split.  A PURE function
// of the container (rows, width, values, blocks) and
// of the process-wide route knobs -- never of free memory, of the self-test of the device at hand or of what happens
// to be resident: the shards of one query run on different devices and partition the pair set differently per
// route (tile raster / row groups / rows), so every shard has to arrive at the same answer.  Whether the chosen
// route can run here (memory) is checked afterwards: an unsharded query may then fall back, a sharded one fails.
int choose_route(const StormState* st, uint64_t n_rows) {
    const bool split_applies = g_sparse_flat && g_sparse_stream && st->n_heavy > 0 && st->n_light >= 2;
    if (g_storm_route == 1) return 1;
    if (g_storm_route == 3) return split_applies ? 4 : 1;
    const uint64_t W = ((uint64_t)st->max_blk_id + 1) * BLOCK_WORDS;
    const bool dense_applies = W < (1u << 25);                           // per-pair counts must stay below 2^31
    if (g_storm_route == 2 && dense_applies) return 2;
    double dense_s = 0, sparse_s = 0;
    storm_route_model(n_rows, W, (double)st->total_nnz / (double)n_rows, (double)st->total_blocks / (double)n_rows,
                      g_sparse_flat && g_sparse_stream && st->n_bitmap_blocks == 0 && st->max_row_nnz <= STREAM_ENTRIES,
                      W * 64 <= (1ull << 24), false, &dense_s, &sparse_s);
    int route = dense_applies && dense_s < sparse_s ? 2 : 1;
    if (split_applies) {
        const double split_s = split_seconds(n_rows, st->n_heavy, (double)st->light_nnz, (double)st->total_nnz, (double)st->max_blocks,
                                             (double)st->n_bitmap_blocks, st->max_blk_id + 1);
        if (split_s < (route == 2 ? dense_s : sparse_s)) route = 4;
    }
    return route;
}

// Whole-container totals on the sparse side read the light mirror when the container has heavy rows (split route) or its
// rows need more than one position range; with one range and no heavy rows the plain flat form serves.
bool route_reads_light(int route, const StormState* st) {
    if (route == 4) return true;
    return route == 1 && st->n_heavy == 0 && st->n_ranges > 1 && g_sparse_flat && g_sparse_stream &&
           st->n_bitmap_blocks == 0 && st->max_row_nnz <= STREAM_ENTRIES;
}

// Does the dense form of the rows fit on this device (keeping 20 % of the free memory)?
bool dense_fits(const StormState* st, uint64_t n_rows) {
    const uint64_t W = ((uint64_t)st->max_blk_id + 1) * BLOCK_WORDS;
    const uint64_t need = n_rows * W * 8;
    if (need <= st->dense_cap_words * 8) return true;                    // (cudaMemGetInfo costs ~0.1 ms: only when the arena must grow)
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return false; }
    const uint64_t have = free_b + (st->d_dense ? st->dense_cap_words * 8 : 0);
    return need <= have / 10 * 8;
}

int ensure_dense(StormState* st, uint64_t n_rows, uint64_t* stride_out) {
    const uint64_t stride = ((uint64_t)st->max_blk_id + 1) * BLOCK_WORDS;   // a multiple of 16 words
    *stride_out = stride;
    if (st->dense_valid) return STORM_B200_OK;
    const uint64_t words = n_rows * stride;
    if (words > st->dense_cap_words) {
        if (st->d_dense) cudaFree(st->d_dense);
        st->d_dense = nullptr; st->dense_cap_words = 0;
        if (cudaMalloc(&st->d_dense, words * 8) != cudaSuccess) {
            cudaGetLastError();
            set_error("dense form of the STORM_t rows (%llu bytes) does not fit", (unsigned long long)(words * 8));
            return STORM_B200_ENOMEM;
        }
        st->dense_cap_words = words;
    }
    STORM_CUDA_TRY(cudaMemsetAsync(st->d_dense, 0, words * 8, st->stream));
    densify_rows_kernel<<<(unsigned)n_rows, 256, 0, st->stream>>>(view_of(st), st->d_dense, stride, 0u);
    STORM_CUDA_TRY(cudaGetLastError());
    count_launch();
    st->dense_valid = true;
    return STORM_B200_OK;
}

// Dense route for containers whose dense form is too large to keep whole (N x W x 8 bytes above DENSE_WHOLE_MAX): the
// rows are densified one band of DENSE_BAND_BYTES at a time into two arenas and the triangle is walked band pair by
// band pair -- the triangle of a band, then its rectangle with every later band -- through the same tile kernels.  The
// merge/probe block kernel (10-70 x slower on such rows) is no longer what a large container falls back to.  Band size
// and the whole / banded decision are fixed byte counts, not functions of free memory: every shard of a sharded query
// cuts the same bands, and shard s takes the band pairs k with k mod n_shards == s.
constexpr uint64_t DENSE_WHOLE_MAX = 48ull << 30;
constexpr uint64_t DENSE_BAND_BYTES = 12ull << 30;

int dense_banded(StormState* st, uint64_t n_rows, uint32_t shard, uint32_t n_shards, uint64_t band_rows_override = 0) {
    const uint64_t stride = ((uint64_t)st->max_blk_id + 1) * BLOCK_WORDS;
    uint64_t R = band_rows_override ? band_rows_override : std::max<uint64_t>(256, DENSE_BAND_BYTES / (stride * 8) / 256 * 256);
    R = std::min(R, n_rows);
    if (R * stride > st->band_cap_words) {
        for (auto& a : st->d_band) { if (a) cudaFree(a); a = nullptr; }
        st->band_cap_words = 0;
        for (auto& a : st->d_band)
            if (cudaMalloc(&a, R * stride * 8) != cudaSuccess) {
                cudaGetLastError(); a = nullptr;
                set_error("two dense bands of %llu rows x %llu words do not fit on device %d", (unsigned long long)R, (unsigned long long)stride, st->device);
                return STORM_B200_ENOMEM;
            }
        st->band_cap_words = R * stride;
    }
    uint64_t* const* arena = st->d_band;                                  // (kept for the next query: everything here is asynchronous)
    auto densify = [&](uint64_t* dst, uint64_t r0, uint64_t n) -> int {
        STORM_CUDA_TRY(cudaMemsetAsync(dst, 0, n * stride * 8, st->stream));
        densify_rows_kernel<<<(unsigned)n, 256, 0, st->stream>>>(view_of(st), dst, stride, (uint32_t)r0);
        STORM_CUDA_TRY(cudaGetLastError());
        count_launch();
        return STORM_B200_OK;
    };
    const uint64_t n_bands = (n_rows + R - 1) / R;
    uint64_t k = 0;                                                        // running band-pair index (I <= J, I-major)
    for (uint64_t I = 0; I < n_bands; ++I) {
        const uint64_t i0 = I * R, ni = std::min(R, n_rows - i0);
        bool have_i = false;
        for (uint64_t J = I; J < n_bands; ++J, ++k) {
            if (k % n_shards != shard) continue;
            int rc = STORM_B200_OK;
            if (!have_i) { if ((rc = densify(arena[0], i0, ni))) return rc; have_i = true; }
            if (J == I) {
                rc = pairw_triangle(arena[0], ni, (uint32_t)stride, stride, 0, 1, STORM_B200_KERNEL_AUTO,
                                    reinterpret_cast<uint64_t*>(st->d_total), st->stream);
            } else {
                const uint64_t j0 = J * R, nj = std::min(R, n_rows - j0);
                if ((rc = densify(arena[1], j0, nj))) return rc;
                rc = pairw_rect(arena[0], ni, stride, i0, arena[1], nj, stride, j0, (uint32_t)stride, 0, STORM_B200_KERNEL_AUTO,
                                nullptr, 0, reinterpret_cast<uint64_t*>(st->d_total), st->stream);
            }
            if (rc) return rc;
        }
    }
    return STORM_B200_OK;
}

std::atomic<uint64_t> g_dense_band_rows{0};   // STORM_b200_set_storm_band_rows: force the banded form with this band height (tests)

// One replica's share of a whole-container query (asynchronous on its stream, accumulated into its d_total).
int storm_query_on(StormState* st, uint32_t n_conts, int route, uint32_t shard, uint32_t n_shards) {
    bool dense = route == 2;
    if (cudaMemsetAsync(st->d_total, 0, 8, st->stream) != cudaSuccess) { set_error("memset failed: %s", cudaGetErrorString(cudaGetLastError())); return STORM_B200_ECUDA; }
    uint64_t stride = 0;
    const uint64_t forced_band = g_dense_band_rows.load();
    if (dense && (forced_band || (uint64_t)n_conts * (((uint64_t)st->max_blk_id + 1) * BLOCK_WORDS) * 8 > DENSE_WHOLE_MAX)) {
        st->last_route = 3;                                              // dense, banded
        return dense_banded(st, n_conts, shard, n_shards, forced_band);
    }
    if (dense && (!dense_fits(st, n_conts) || ensure_dense(st, n_conts, &stride) != STORM_B200_OK)) {
        // the dense form does not fit on this device: an unsharded query answers through the sparse kernels instead;
        // a shard must not (the other shards partition the pairs by the tile raster)
        if (n_shards > 1) { set_error("shard %u of %u: the dense form of the rows does not fit on device %d and a shard cannot switch route", shard, n_shards, st->device); return STORM_B200_ENOMEM; }
        dense = false;
    }
    if (dense) {
        st->last_route = 2;
        return pairw_triangle(st->d_dense, n_conts, (uint32_t)stride, stride, shard, n_shards, STORM_B200_KERNEL_AUTO,
                              reinterpret_cast<uint64_t*>(st->d_total), st->stream);
    }
    int rc = STORM_B200_OK;
    // The light rows through the stream kernel, range by range of the light mirror (one range and no heavy rows: the
    // plain flat form below); then, on the split route, every pair with a heavy row through the block kernel.
    const bool ranged = st->light_mirror && route_reads_light(route, st);
    if (ranged) {
        bool ok = false;
        if ((rc = ensure_light_flat(st, &ok))) return rc;
        if (ok) {
            st->last_route = route == 4 ? 4 : 1;
            const uint64_t stride = (uint64_t)st->n_light + 1;
            for (uint32_t q = 0; q < st->n_ranges; ++q)
                if ((rc = launch_sparse_stream(st->d_lpos_off + q * stride, st->d_lpos, st->h_lgroup_start, st->d_lgroup_start,
                                               st->d_lpos_off + q * stride, st->d_lpos, st->range_nnz[q], 0, st->n_light, 0, st->n_light, 1,
                                               shard, n_shards, st->d_total, st->stream)))
                    return rc;
            if (route != 4) return STORM_B200_OK;
            SparseJob job{};
            job.A = job.B = view_of(st);
            job.i_list = st->d_heavy_rows;
            job.i0 = 0; job.i1 = st->n_heavy; job.j0 = 0; job.j1 = n_conts;
            job.shard = shard; job.n_shards = n_shards;
            job.total = st->d_total;
            return launch_sparse(job, st->max_blocks, st->stream);
        }
        if (n_shards > 1) { set_error("shard %u of %u: no room for the light rows' position mirror and a shard cannot switch kernel", shard, n_shards); return STORM_B200_ENOMEM; }
    }
    st->last_route = 1;
    if ((rc = ensure_flat(st))) return rc;
    // (same rule inside the sparse route: the stream kernel shards row groups, the other two rows)
    if (n_shards > 1 && flat_eligible(st) && !st->flat_valid) { set_error("shard %u of %u: no room for the flat position mirror and a shard cannot switch kernel", shard, n_shards); return STORM_B200_ENOMEM; }
    if (stream_eligible(st, st)) return launch_stream(st, st, 0, n_conts, 0, n_conts, 1, shard, n_shards, st->d_total, st->stream);
    SparseJob job{};
    job.A = job.B = view_of(st);
    job.i0 = 0; job.i1 = n_conts; job.j0 = 0; job.j1 = n_conts;
    job.strict_upper = 1;
    job.shard = shard; job.n_shards = n_shards;
    job.total = st->d_total;
    return launch_sparse(job, st->max_blocks, st->stream);
}

// Whole-container query: replica g of G (the device set, devices.h) answers shard (shard * G + g) of (n_shards * G)
// on its own device and stream; the host adds the G totals.  The route is a pure function of the container, so every
// replica -- and every shard of a multi-process caller -- takes the same one.
uint64_t storm_query(STORM_t* s, uint32_t shard, uint32_t n_shards) {
    if (s == nullptr) return (uint64_t)-1;                          // storm.c:878,898
    if (s->n_conts < 2) return 0;
    StormState* st = state_of(s);
    DeviceGuard home(st->device);
    if (sync_mirror(s, st, true)) return (uint64_t)-1;
    std::vector<StormState*> reps{st};
    reps.insert(reps.end(), st->replicas.begin(), st->replicas.end());
    const uint32_t G = (uint32_t)reps.size();
    const int route = choose_route(st, s->n_conts);
    if (route_reads_light(route, st) && sync_light(s, st)) return (uint64_t)-1;
    // one host thread per replica (devices.h: for_each_device): launch its share, read its total back, wait
    const int rc = for_each_device((int)G, [&](int g) -> int {
        StormState* r = reps[g];
        DeviceGuard guard(r->device);
        int qrc = storm_query_on(r, s->n_conts, route, shard * G + (uint32_t)g, n_shards * G);
        if (qrc) return qrc;
        if (cudaMemcpyAsync(r->h_total, r->d_total, 8, cudaMemcpyDeviceToHost, r->stream) != cudaSuccess ||
            cudaStreamSynchronize(r->stream) != cudaSuccess) {
            set_error("STORM_t query failed on device %d: %s", r->device, cudaGetErrorString(cudaGetLastError()));
            return STORM_B200_ECUDA;
        }
        return STORM_B200_OK;
    });
    if (rc) return (uint64_t)-1;
    uint64_t total = 0;
    for (uint32_t g = 0; g < G; ++g) total += *reps[g]->h_total;
    return total;
}

void mark_dirty(STORM_t* s) {
    if (!s || !s->b200) return;
    StormState* st = state_of(s);
    st->dirty = true;
    for (StormState* r : st->replicas) r->dirty = true;
}

void release_state(StormState* st) {
    DeviceGuard guard(st->device);
    if (st->stream) cudaStreamSynchronize(st->stream);
    free_mirror(st);
    for (auto& a : st->d_band) if (a) cudaFree(a);
    if (st->d_dense) cudaFree(st->d_dense);
    if (st->d_total) cudaFree(st->d_total);
    if (st->h_total) cudaFreeHost(st->h_total);
    if (st->stream) cudaStreamDestroy(st->stream);
}

}  // namespace
}  // namespace storm

// =================================================================================
// C ABI: host containers
// =================================================================================
using namespace storm;

extern "C" {

// ---- block (storm.c:398-569) -------------------------------------------------------
void STORM_bitmap_init(STORM_bitmap_t* b) {                                // storm.c:416-430
    if (b == nullptr) return;
    memset(b, 0, sizeof(*b));
    b->own_data = 1;
    b->own_scalar = 1;
}

STORM_bitmap_t* STORM_bitmap_new(void) {                                    // storm.c:398-413
    void* p = nullptr;
    if (posix_memalign(&p, 64, sizeof(STORM_bitmap_t))) return nullptr;
    STORM_bitmap_init((STORM_bitmap_t*)p);
    return (STORM_bitmap_t*)p;
}

static void bitmap_release(STORM_bitmap_t* b) {
    if (b->own_data) free(b->data);
    if (b->own_scalar) free(b->scalar);
    b->data = nullptr; b->scalar = nullptr;
}

void STORM_bitmap_free(STORM_bitmap_t* b) {                                 // storm.c:433-438
    if (b == nullptr) return;
    bitmap_release(b);
    free(b);
}

int STORM_bitmap_add(STORM_bitmap_t* b, const uint32_t* values, const uint32_t n_values) {   // storm.c:442-465
    if (b == nullptr) return -1;
    if (values == nullptr) return -2;
    if (n_values == 0) return -3;
    const uint32_t adjust = b->id * BLOCK_BITS;
    if (b->data == nullptr) {
        void* p = nullptr;
        if (posix_memalign(&p, 64, BLOCK_WORDS * sizeof(uint64_t))) return -4;
        memset(p, 0, BLOCK_WORDS * sizeof(uint64_t));
        b->data = (uint64_t*)p;
    }
    b->n_bitmap = BLOCK_WORDS;
    for (uint32_t i = 0; i < n_values; ++i) {
        const uint32_t v = values[i] - adjust;
        if (v >= BLOCK_BITS) return -5;                                   // reference: assert, compiled out
        const uint64_t bit = 1ull << (v & 63);
        b->n_bits_set += (b->data[v >> 6] & bit) == 0;
        b->data[v >> 6] |= bit;
    }
    return (int)n_values;
}

int STORM_bitmap_add_scalar_only(STORM_bitmap_t* b, const uint32_t* values, const uint32_t n_values) {  // :521-558
    if (b == nullptr) return -1;
    if (values == nullptr) return -3;
    if (n_values == 0) return -4;
    const uint32_t need = b->n_scalar + n_values;
    if (b->scalar == nullptr || need > b->m_scalar) {                     // capacity checked against the need (D9)
        const uint32_t cap = std::max<uint32_t>(256, need + (need >> 2));
        uint16_t* p = (uint16_t*)realloc(b->own_scalar ? b->scalar : nullptr, cap * sizeof(uint16_t));
        if (p == nullptr) return -5;
        b->scalar = p; b->m_scalar = cap; b->own_scalar = 1;
    }
    const uint32_t adjust = b->id * BLOCK_BITS;
    b->n_scalar_set = 1;
    uint32_t n = b->n_scalar;
    for (uint32_t i = 0; i < n_values; ++i) {
        const uint32_t v = values[i] - adjust;
        if (v >= BLOCK_BITS) return -5;
        b->scalar[n++] = (uint16_t)v;
        ++b->n_bits_set;
    }
    b->n_scalar = n;
    return (int)n_values;
}

int STORM_bitmap_clear(STORM_bitmap_t* b) {                                 // storm.c:561-569
    if (b == nullptr) return -1;
    if (b->data != nullptr) memset(b->data, 0, sizeof(uint64_t) * BLOCK_WORDS);
    b->n_scalar = 0;
    b->n_bits_set = 0;
    b->n_bitmap = 0;
    return 1;
}

uint32_t STORM_bitmap_serialized_size(STORM_bitmap_t* b) {                  // storm.c:372-381
    uint32_t total = (uint32_t)sizeof(uint64_t) * b->n_bitmap;
    if (b->n_scalar_set) total += (uint32_t)sizeof(uint16_t) * b->n_scalar;
    return total + 4 * (uint32_t)sizeof(uint32_t);
}

// ---- row (storm.c:659-824) ---------------------------------------------------------
void STORM_bitmap_cont_init(STORM_bitmap_cont_t* r) {                       // storm.c:670-677
    if (r == nullptr) return;
    memset(r, 0, sizeof(*r));
}

STORM_bitmap_cont_t* STORM_bitmap_cont_new(void) {                          // storm.c:659-668
    STORM_bitmap_cont_t* r = (STORM_bitmap_cont_t*)malloc(sizeof(STORM_bitmap_cont_t));
    STORM_bitmap_cont_init(r);
    return r;
}

static void cont_release(STORM_bitmap_cont_t* r) {
    for (uint32_t i = 0; i < r->m_bitmaps; ++i) bitmap_release(&r->bitmaps[i]);
    free(r->bitmaps);
    free(r->block_ids);
    r->bitmaps = nullptr; r->block_ids = nullptr; r->n_bitmaps = r->m_bitmaps = 0;
}

void STORM_bitmap_cont_free(STORM_bitmap_cont_t* r) {                       // storm.c:679-689
    if (r == nullptr) return;
    cont_release(r);
    free(r);
}

static int cont_reserve(STORM_bitmap_cont_t* r, uint32_t need) {
    if (need <= r->m_bitmaps) return 0;
    const uint32_t cap = std::max<uint32_t>(need, r->m_bitmaps ? r->m_bitmaps + 8 : 2);   // :699,:728-735
    void* p = nullptr;                                                     // STORM_bitmap_t is 64-byte aligned
    if (posix_memalign(&p, 64, cap * sizeof(STORM_bitmap_t))) return -1;
    if (r->bitmaps) memcpy(p, r->bitmaps, r->m_bitmaps * sizeof(STORM_bitmap_t));
    free(r->bitmaps);
    r->bitmaps = (STORM_bitmap_t*)p;
    for (uint32_t i = r->m_bitmaps; i < cap; ++i) STORM_bitmap_init(&r->bitmaps[i]);
    uint32_t* ids = (uint32_t*)realloc(r->block_ids, cap * sizeof(uint32_t));
    if (ids == nullptr) return -1;
    r->block_ids = ids;
    r->m_bitmaps = cap;
    return 0;
}

int STORM_bitmap_cont_add(STORM_bitmap_cont_t* r, const uint32_t* values, const uint32_t n_values) {  // :692-758
    if (r == nullptr) return -1;
    if (values == nullptr) return -2;
    if (n_values == 0) return 0;
    uint32_t start = 0;
    while (start < n_values) {
        const uint32_t target = values[start] / BLOCK_BITS;
        uint32_t stop = start;
        while (stop < n_values && values[stop] / BLOCK_BITS == target) ++stop;
        if (cont_reserve(r, r->n_bitmaps + 1)) return -3;
        STORM_bitmap_t* x = &r->bitmaps[r->n_bitmaps];
        x->id = target;
        r->block_ids[r->n_bitmaps] = target;
        int rc;
        if (stop - start < LIST_THRESHOLD) rc = STORM_bitmap_add_scalar_only(x, values + start, stop - start);  // :745-746
        else rc = STORM_bitmap_add(x, values + start, stop - start);                                          // :747-748
        if (rc < 0) return -3;
        ++r->n_bitmaps;
        r->prev_inserted_value = values[stop - 1];
        start = stop;
    }
    return 1;
}

int STORM_bitmap_cont_clear(STORM_bitmap_cont_t* r) {                       // storm.c:816-824
    if (r == nullptr) return -1;
    for (uint32_t i = 0; i < r->n_bitmaps; ++i) STORM_bitmap_clear(&r->bitmaps[i]);
    r->n_bitmaps = 0;
    r->prev_inserted_value = 0;
    return 1;
}

uint32_t STORM_bitmap_cont_serialized_size(STORM_bitmap_cont_t* r) {        // storm.c:384-394
    uint32_t total = 0;
    if (r->bitmaps != nullptr)
        for (uint32_t i = 0; i < r->n_bitmaps; ++i) total += STORM_bitmap_serialized_size(&r->bitmaps[i]);
    return total + (uint32_t)sizeof(uint32_t) * r->n_bitmaps + 3 * (uint32_t)sizeof(uint32_t);
}

// ---- top (storm.c:827-973) ---------------------------------------------------------
STORM_t* STORM_new(void) {                                                  // storm.c:827-834
    STORM_t* s = (STORM_t*)calloc(1, sizeof(STORM_t));
    if (s == nullptr) return nullptr;
    StormState* st = new (std::nothrow) StormState();
    if (st == nullptr) { free(s); return nullptr; }
    s->b200 = st;
    int n = 0;
    if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) cudaGetDevice(&st->device); else cudaGetLastError();
    return s;
}

void STORM_free(STORM_t* s) {                                               // storm.c:836-842 (+ D8)
    if (s == nullptr) return;
    StormState* st = state_of(s);
    if (st) {
        for (StormState* r : st->replicas) { release_state(r); delete r; }
        release_state(st);
        delete st;
    }
    for (uint32_t i = 0; i < s->m_conts; ++i) cont_release(&s->conts[i]);
    free(s->conts);
    free(s);
}

int STORM_add(STORM_t* s, const uint32_t* values, const uint32_t n_values) {   // storm.c:844-866
    if (s == nullptr) return -1;
    if (s->n_conts == s->m_conts) {
        const uint32_t cap = s->m_conts + 1024;
        STORM_bitmap_cont_t* c = (STORM_bitmap_cont_t*)realloc(s->conts, cap * sizeof(STORM_bitmap_cont_t));
        if (c == nullptr) return -3;
        for (uint32_t i = s->m_conts; i < cap; ++i) STORM_bitmap_cont_init(&c[i]);
        s->conts = c; s->m_conts = cap;
    }
    STORM_bitmap_cont_add(&s->conts[s->n_conts++], values, n_values);      // an empty list still appends a row
    mark_dirty(s);
    return 1;
}

int STORM_clear(STORM_t* s) {                                               // storm.c:868-875
    if (s == nullptr) return -1;
    for (uint32_t i = 0; i < s->n_conts; ++i) STORM_bitmap_cont_clear(&s->conts[i]);
    s->n_conts = 0;
    mark_dirty(s);
    return 1;
}

uint64_t STORM_serialized_size(const STORM_t* s) {                          // storm.c:963-973
    if (s == nullptr) return 0;
    uint64_t tot = 0;
    for (uint32_t i = 0; i < s->n_conts; ++i) tot += STORM_bitmap_cont_serialized_size(&s->conts[i]);
    return tot + 2 * sizeof(uint32_t);
}

uint64_t STORM_pairw_intersect_cardinality(STORM_t* s) {                    // storm.c:877-895
    return storm_query(s, 0, 1);
}

uint64_t STORM_pairw_intersect_cardinality_blocked(STORM_t* s, uint32_t bsize) {   // storm.c:897-961
    (void)bsize;   // 0 = auto on the CPU (storm.c:903-914); a hint only, the total does not depend on it
    return storm_query(s, 0, 1);
}

uint64_t STORM_b200_storm_pairw_shard(STORM_t* s, uint32_t shard, uint32_t n_shards) {
    if (s == nullptr) return (uint64_t)-1;
    if (n_shards == 0 || shard >= n_shards) { set_error("shard %u of %u", shard, n_shards); return (uint64_t)-1; }
    return storm_query(s, shard, n_shards);
}

int STORM_b200_set_storm_route(int route) {
    const int prev = g_storm_route;
    if (route >= 0 && route <= 3) g_storm_route = route;
    return prev;
}

// Seconds the route model expects for a whole-container STORM_t query: out[0] densified rows + tile kernel, out[1]
// sparse kernels.  Pure arithmetic (no device needed); the query takes the smaller one.
int STORM_b200_storm_route_model(uint64_t n_rows, uint32_t n_words, double avg_nnz, double avg_blocks, uint32_t max_row_nnz,
                                 uint64_t n_bitmap_blocks, int fp4, int dense_resident, double out_seconds[2]) {
    if (out_seconds == nullptr || n_rows < 2 || n_words == 0) { set_error("bad argument"); return STORM_B200_EINVAL; }
    storm_route_model(n_rows, n_words, avg_nnz, avg_blocks, n_bitmap_blocks == 0 && max_row_nnz <= STREAM_ENTRIES, fp4 != 0,
                      dense_resident != 0, &out_seconds[0], &out_seconds[1]);
    return STORM_B200_OK;
}

// Seconds the model expects for the split route (light rows: stream kernel; pairs with a heavy row: block kernel).
double STORM_b200_storm_split_model(uint64_t n_rows, uint64_t n_heavy, double light_nnz, double total_nnz, double max_blocks,
                                    double n_bitmap_blocks) {
    if (n_rows < 2 || n_heavy > n_rows) return -1.0;
    return split_seconds(n_rows, n_heavy, light_nnz, total_nnz, max_blocks, n_bitmap_blocks, (uint32_t)std::max(1.0, max_blocks));
}

int STORM_b200_set_sparse_flat(int mode) {
    const int prev = g_sparse_flat ? (g_sparse_stream ? 2 : 1) : 0;
    if (mode >= 0 && mode <= 2) { g_sparse_flat = mode >= 1; g_sparse_stream = mode == 2; }
    return prev;
}

// Tests / measurements: force the banded dense form with bands of `rows` rows (0 = back to the size rule).  Returns the previous value.
uint64_t STORM_b200_set_storm_band_rows(uint64_t rows) { return g_dense_band_rows.exchange(rows); }

int STORM_b200_storm_last_route(const STORM_t* s) {
    if (s == nullptr || s->b200 == nullptr) return 0;
    return state_of(s)->last_route;
}

int STORM_b200_storm_pairw_rect(STORM_t* s, uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, uint32_t* out) {
    if (s == nullptr || out == nullptr) { set_error("NULL argument"); return STORM_B200_EINVAL; }
    if (i0 > i1 || j0 > j1 || i1 > s->n_conts || j1 > s->n_conts) { set_error("rectangle outside the %u rows", s->n_conts); return STORM_B200_EINVAL; }
    if (i0 == i1 || j0 == j1) return STORM_B200_OK;
    StormState* st = state_of(s);
    DeviceGuard guard(st->device);
    int rc = sync_mirror(s, st);
    if (rc) return rc;
    const uint64_t ni = i1 - i0, nj = j1 - j0;
    uint32_t* d_out = nullptr;
    if (cudaMalloc(&d_out, ni * nj * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); set_error("device allocation for %llu x %llu counts failed", (unsigned long long)ni, (unsigned long long)nj); return STORM_B200_ENOMEM; }
    // Containers with bitmap blocks: the two row ranges are densified and the rectangle goes through the tile kernel
    // (per-pair form) -- the block merge/probe kernel is 10-70 x slower on such rows.  Containers of list blocks keep
    // the flat probe kernel below.
    const uint64_t dense_stride = ((uint64_t)st->max_blk_id + 1) * BLOCK_WORDS;
    if (st->n_bitmap_blocks > 0 && g_storm_route != 1 && dense_stride < (1u << 25) && (ni + nj) * dense_stride * 8 <= DENSE_WHOLE_MAX / 2) {
        uint64_t* d_rows = nullptr;
        if (cudaMalloc(&d_rows, (ni + nj) * dense_stride * 8) == cudaSuccess) {
            rc = STORM_B200_OK;
            if (cudaMemsetAsync(d_rows, 0, (ni + nj) * dense_stride * 8, st->stream) != cudaSuccess) rc = STORM_B200_ECUDA;
            if (!rc) {
                densify_rows_kernel<<<(unsigned)ni, 256, 0, st->stream>>>(view_of(st), d_rows, dense_stride, (uint32_t)i0);
                densify_rows_kernel<<<(unsigned)nj, 256, 0, st->stream>>>(view_of(st), d_rows + ni * dense_stride, dense_stride, (uint32_t)j0);
                count_launch(2);
                if (cudaGetLastError() != cudaSuccess) rc = STORM_B200_ECUDA;
            }
            if (!rc) rc = pairw_rect(d_rows, ni, dense_stride, i0, d_rows + ni * dense_stride, nj, dense_stride, j0, (uint32_t)dense_stride, 1,
                                     STORM_B200_KERNEL_AUTO, d_out, nj, nullptr, st->stream);
            if (!rc && (cudaMemcpyAsync(out, d_out, ni * nj * sizeof(uint32_t), cudaMemcpyDeviceToHost, st->stream) != cudaSuccess ||
                        cudaStreamSynchronize(st->stream) != cudaSuccess)) {
                set_error("STORM_t rect query failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = STORM_B200_ECUDA;
            }
            cudaFree(d_rows); cudaFree(d_out);
            return rc;
        }
        cudaGetLastError();                                              // no room for the dense rows: the block kernel answers
    }
    if ((rc = ensure_flat(st))) { cudaFree(d_out); return rc; }
    SparseJob job{};
    job.A = job.B = view_of(st);
    job.i0 = i0; job.i1 = i1; job.j0 = j0; job.j1 = j1;
    job.strict_upper = 1; job.shard = 0; job.n_shards = 1;
    job.out = d_out; job.ld = nj;
    if (cudaMemsetAsync(d_out, 0, ni * nj * sizeof(uint32_t), st->stream) != cudaSuccess) rc = STORM_B200_ECUDA;
    if (!rc) rc = launch_sparse(job, st->max_blocks, st->stream);
    if (!rc && (cudaMemcpyAsync(out, d_out, ni * nj * sizeof(uint32_t), cudaMemcpyDeviceToHost, st->stream) != cudaSuccess ||
                cudaStreamSynchronize(st->stream) != cudaSuccess)) {
        set_error("STORM_t rect query failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = STORM_B200_ECUDA;
    }
    cudaFree(d_out);
    return rc;
}

// Declared in the reference (storm.h:229) but never defined there (storm.c:975).
uint64_t STORM_intersect_cardinality_square(const STORM_t* STORM_RESTRICT s1, const STORM_t* STORM_RESTRICT s2) {
    if (s1 == nullptr || s2 == nullptr) return (uint64_t)-1;
    if (s1->n_conts == 0 || s2->n_conts == 0) return 0;
    StormState* a = state_of(s1);
    StormState* b = state_of(s2);
    DeviceGuard guard(a->device);
    if (sync_mirror(s1, a) || sync_mirror(s2, b)) return (uint64_t)-1;
    if (a->device != b->device) { set_error("the two containers live on different devices"); return (uint64_t)-1; }
    if (cudaStreamSynchronize(b->stream) != cudaSuccess) return (uint64_t)-1;
    if (cudaMemsetAsync(a->d_total, 0, 8, a->stream) != cudaSuccess) return (uint64_t)-1;
    if (flat_eligible(a) && flat_eligible(b) && (ensure_flat(a) || ensure_flat(b))) return (uint64_t)-1;
    if (cudaStreamSynchronize(b->stream) != cudaSuccess) return (uint64_t)-1;   // b's flat form is built on b's stream
    if (stream_eligible(a, b)) {
        if (launch_stream(a, b, 0, s1->n_conts, 0, s2->n_conts, 0, 0, 1, a->d_total, a->stream)) return (uint64_t)-1;
    } else {
        SparseJob job{};
        job.A = view_of(a); job.B = view_of(b);
        job.i0 = 0; job.i1 = s1->n_conts; job.j0 = 0; job.j1 = s2->n_conts;
        job.strict_upper = 0; job.shard = 0; job.n_shards = 1;
        job.total = a->d_total;
        if (launch_sparse(job, a->max_blocks, a->stream)) return (uint64_t)-1;
    }
    if (cudaMemcpyAsync(a->h_total, a->d_total, 8, cudaMemcpyDeviceToHost, a->stream) != cudaSuccess ||
        cudaStreamSynchronize(a->stream) != cudaSuccess) {
        set_error("square query failed: %s", cudaGetErrorString(cudaGetLastError()));
        return (uint64_t)-1;
    }
    return *a->h_total;
}

}  // extern "C"
