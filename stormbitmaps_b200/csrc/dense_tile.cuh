// dense_tile.cuh -- the CUDA-core tile kernel shared by dense_popc.cu and dense_csa.cu.
//
// Replaces the reference's hot loop `count += f(row_i, row_j, W)` over all
// pairs (storm.c:1165-1169, blocked form :1199-1238) and its per-pair SIMD
// kernel (libalgebra.h:2684-2744, 2872-2890) for one 128 x 128 tile of pairs
// per CTA.
//
// Shape of the computation:
//   * CTA = 256 threads = 16 x 16 grid; thread (ty, tx) owns the 8 x 8 pairs
//     {A rows ty + 16 r} x {B rows tx + 16 c}, one uint32 accumulator each
//     (a pair count is at most M < 2^32).
//   * K (the words of a row) is walked in slabs of 16 words; a slab of both
//     operands (2 x 128 rows x 128 B = 32 KiB) is staged in shared memory by
//     cp.async in 16-byte chunks, 3 stages deep.  Shared layout is
//     [chunk][row][16 B] so that a warp's B reads are 16 consecutive 16-byte
//     chunks (conflict-free) and its A reads are 2 addresses (broadcast).
//   * The inner product over one slab is the template parameter:
//       DIRECT  2 LOP3 + 2 POPC per 64-bit word pair -- bound by the POPC pipe;
//       CSA     a 7:3 carry-save compressor over seven of every eight 32-bit words
//               (the register-level form of the reference's Harley-Seal loop,
//               libalgebra.h:2287-2292 / 2704-2722): 16 LOP3 + 4 POPC per eight
//               words, which balances the ALU and POPC pipes.
//   * Epilogue: optional per-pair store, diagonal mask (global j > i), warp
//     shuffle + shared reduction, ONE 64-bit atomicAdd per CTA.
#pragma once

#include "common.cuh"

namespace storm {
namespace tile {

constexpr int TM = 128, TN = 128;        // pairs per CTA: TM A-rows x TN B-rows
constexpr int THREADS = 256;
constexpr int CHUNKS = 8;                // 16-byte chunks per row per slab (= 16 words)
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = (TM + TN) * CHUNKS * 16;   // 32 KiB

enum Inner { DIRECT = 0, CSA = 1 };

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Stage one slab: rows of A then rows of B, chunk-major.  Lanes walk rows so the
// shared-memory side is contiguous; the global side is 16 B per row (the slab is
// 32 KiB per 262144 word pairs -- global efficiency is irrelevant here).
__device__ __forceinline__ void load_slab(const DenseJob& job, uint32_t smem_stage, uint64_t rowA0,
                                          uint64_t rowB0, uint32_t slab, uint32_t n_chunks_total) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int it = 0; it < (TM + TN) * CHUNKS / THREADS; ++it) {
        const int idx = it * THREADS + tid;          // 0 .. 2047
        const int row = idx & (TM + TN - 1);         // 0..255: A rows then B rows
        const int ch = idx >> 8;                     // 0..7
        const uint32_t gchunk = slab * CHUNKS + ch;
        const bool isB = row >= TM;
        const uint64_t r = isB ? rowB0 + (row - TM) : rowA0 + row;
        const uint64_t nrows = isB ? job.nB : job.nA;
        const uint64_t* base = isB ? job.B : job.A;
        const uint64_t stride = isB ? job.strideB : job.strideA;
        uint32_t bytes = 0;
        if (r < nrows && gchunk < n_chunks_total) {
            const uint32_t words_left = job.n_words - gchunk * 2;
            bytes = words_left >= 2 ? 16u : 8u;
        }
        const uint64_t* src = bytes ? base + r * stride + (uint64_t)gchunk * 2 : base;
        cp_async16_zfill(smem_stage + (uint32_t)(ch * (TM + TN) + row) * 16u, src, bytes);
    }
}

// carry-save adder: (sum, carry) of three bit vectors, one LOP3 each.  Written as PTX so that
// ptxas keeps the 2-instruction form; left to itself it folds the ANDs that produce x, y, z
// into the XOR chain and then needs 22 LOP3 per eight words instead of 16.
__device__ __forceinline__ void csa(uint32_t& s, uint32_t& c, uint32_t x, uint32_t y, uint32_t z) {
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s) : "r"(x), "r"(y), "r"(z));   // x ^ y ^ z
    asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(c) : "r"(x), "r"(y), "r"(z));   // majority
}
__device__ __forceinline__ uint32_t and2(uint32_t x, uint32_t y) {
    uint32_t r;
    asm("and.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    return r;
}

// popcount of eight 32-bit AND results: seven through a 7:3 compressor, the eighth directly
__device__ __forceinline__ uint32_t and_popc8(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    uint32_t s0, c0, s1, c1, ones, c2, twos, fours;
    csa(s0, c0, and2(a0.x, b0.x), and2(a0.y, b0.y), and2(a0.z, b0.z));
    csa(s1, c1, and2(a0.w, b0.w), and2(a1.x, b1.x), and2(a1.y, b1.y));
    csa(ones, c2, s0, s1, and2(a1.z, b1.z));
    csa(twos, fours, c0, c1, c2);
    // IMAD keeps the weighting on the FMA pipe, off the ALU pipe the LOP3s use
    return __popc(ones) + __popc(a1.w & b1.w) + 2u * __popc(twos) + 4u * __popc(fours);
}

template <int INNER>
__global__ void __launch_bounds__(THREADS, 1) dense_tile_kernel(const DenseJob job) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ unsigned long long warp_part[THREADS / 32];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    uint32_t bi, bj;
    tile_coords(job, job.tile_begin + blockIdx.x, TM, TN, bi, bj);
    const uint64_t rowA0 = (uint64_t)bi * TM, rowB0 = (uint64_t)bj * TN;

    const uint32_t n_chunks_total = (job.n_words + 1) / 2;
    const uint32_t n_slabs = (n_chunks_total + CHUNKS - 1) / CHUNKS;
    const uint32_t smem_base = smem_u32(smem);

    uint32_t acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0;

    // prologue: STAGES-1 slabs in flight
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if ((uint32_t)s < n_slabs) load_slab(job, smem_base + s * STAGE_BYTES, rowA0, rowB0, s, n_chunks_total);
        cp_async_commit();
    }

    for (uint32_t slab = 0; slab < n_slabs; ++slab) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();                           // slab landed for everyone; previous compute done
        {
            const uint32_t nxt = slab + STAGES - 1;
            if (nxt < n_slabs) load_slab(job, smem_base + (nxt % STAGES) * STAGE_BYTES, rowA0, rowB0, nxt, n_chunks_total);
            cp_async_commit();
        }
        const uint4* sA = reinterpret_cast<const uint4*>(smem + (slab % STAGES) * STAGE_BYTES);
        const uint4* sB = sA + TM;
        if (INNER == DIRECT) {
#pragma unroll 1   // one chunk = 256 POPC per thread; unrolling further only spills
            for (int ch = 0; ch < CHUNKS; ++ch) {
                uint4 a[8], b[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) a[r] = sA[ch * (TM + TN) + ty + 16 * r];
#pragma unroll
                for (int c = 0; c < 8; ++c) b[c] = sB[ch * (TM + TN) + tx + 16 * c];
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        acc[r][c] += __popc(a[r].x & b[c].x) + __popc(a[r].y & b[c].y) +
                                     __popc(a[r].z & b[c].z) + __popc(a[r].w & b[c].w);
            }
        } else {
#pragma unroll 1   // two chunks = eight 32-bit words per pair and step
            for (int ch = 0; ch < CHUNKS; ch += 2) {
                uint4 a0[8], a1[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    a0[r] = sA[ch * (TM + TN) + ty + 16 * r];
                    a1[r] = sA[(ch + 1) * (TM + TN) + ty + 16 * r];
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 b0 = sB[ch * (TM + TN) + tx + 16 * c];
                    const uint4 b1 = sB[(ch + 1) * (TM + TN) + tx + 16 * c];
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[r][c] += and_popc8(a0[r], a1[r], b0, b1);
                }
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue -------------------------------------------------------------
    unsigned long long sum = 0;
    const uint64_t gi0 = job.i_off + rowA0, gj0 = job.j_off + rowB0;
    // the tile needs per-element masking only if it touches the diagonal
    const bool diag = job.strict_upper && (gj0 <= gi0 + TM - 1);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint64_t li = rowA0 + ty + 16 * r;
        if (li >= job.nA) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint64_t lj = rowB0 + tx + 16 * c;
            if (lj >= job.nB) continue;
            uint32_t v = acc[r][c];
            if (diag && (job.j_off + lj <= job.i_off + li)) v = 0;
            sum += v;
            if (job.out) job.out[li * job.ld + lj] = v;
        }
    }
    if (job.total) {
        sum = warp_sum(sum);
        if ((tid & 31) == 0) warp_part[tid >> 5] = sum;
        __syncthreads();
        if (tid == 0) {
            unsigned long long t = 0;
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) t += warp_part[w];
            if (t) atomicAdd(job.total, t);
        }
    }
}

template <int INNER>
int launch_dense_tile(const DenseJob& job, cudaStream_t stream) {
    if (job.tile_end <= job.tile_begin) return STORM_B200_OK;
    const int smem_bytes = STAGES * STAGE_BYTES;
    // per device and cheap: set on every launch instead of caching a flag
    STORM_CUDA_TRY(cudaFuncSetAttribute(dense_tile_kernel<INNER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    uint64_t remaining = job.tile_end - job.tile_begin, begin = job.tile_begin;
    while (remaining) {                                // grid.x is limited to 2^31 - 1
        const uint64_t n = remaining > 0x40000000ull ? 0x40000000ull : remaining;
        DenseJob j = job;
        j.tile_begin = begin;
        j.tile_end = begin + n;
        dense_tile_kernel<INNER><<<(unsigned)n, THREADS, smem_bytes, stream>>>(j);
        STORM_CUDA_TRY(cudaGetLastError());
        count_launch();
        begin += n;
        remaining -= n;
    }
    return STORM_B200_OK;
}

}  // namespace tile
}  // namespace storm
