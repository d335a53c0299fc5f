// dense_b1.cu -- the legacy one-bit tensor path, for the record: mma.sync.m16n8k256 .b1 AND + POPC.
//
// The north star names two tensor-core formulations to be benchmarked against LOP3 + POPC: mma.sync b1 AND.popc
// "where sm_100a exposes it" and the bit-unpacked tcgen05 GEMM.  sm_100a accepts the instruction but has no BMMA
// pipe: ptxas lowers one m16n8k256 to IMMA.16832.U8.U8 steps plus ALU glue (SASS excerpt in
// profiles/r02_dense_b1_sass.txt).  This kernel is a straightforward 128 x 128 tile around it -- cp.async slabs,
// swizzled shared memory, eight warps of 32 x 64 pairs -- so that its rate on the box stands beside the other four
// forms (profiles/README.md).  It is exact (s32 accumulators) and selectable as STORM_B200_KERNEL_B1; AUTO never picks it.
//
// Replaces the same loop nest as the other dense kernels: storm.c:1165-1169 / 1199-1238 + libalgebra.h:2872-2890.
#include "dense_tile.cuh"

namespace storm {
namespace b1 {

constexpr int TM = 128, TN = 128;
constexpr int THREADS = 256;                 // 8 warps: 4 along A rows x 2 along B rows, 32 x 64 pairs each
constexpr int SLAB_BYTES = 128;              // 1024 bits of K per row and slab = four k256 steps
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = (TM + TN) * SLAB_BYTES;   // 32 KiB

// D (16 x 8, s32) += popc(A (16 x 256 bits, row) & B (256 bits x 8, col))
__device__ __forceinline__ void mma_b1(uint32_t (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// rows of A then rows of B, [row][128 B], 16-byte chunk c of row r stored at chunk c ^ (r & 7)
__device__ __forceinline__ void load_slab(const DenseJob& job, uint32_t smem_stage, uint64_t rowA0, uint64_t rowB0, uint32_t slab) {
    const int tid = threadIdx.x;
    const uint32_t n_chunks_total = (job.n_words + 1) / 2;
#pragma unroll
    for (int it = 0; it < (TM + TN) * 8 / THREADS; ++it) {
        const int idx = it * THREADS + tid;
        const int row = idx >> 3, ch = idx & 7;
        const uint32_t gchunk = slab * 8 + ch;
        const bool isB = row >= TM;
        const uint64_t r = isB ? rowB0 + (row - TM) : rowA0 + row;
        const uint64_t nrows = isB ? job.nB : job.nA;
        const uint64_t* base = isB ? job.B : job.A;
        const uint64_t stride = isB ? job.strideB : job.strideA;
        uint32_t bytes = 0;
        if (r < nrows && gchunk < n_chunks_total) bytes = job.n_words - gchunk * 2 >= 2 ? 16u : 8u;
        const uint64_t* src = bytes ? base + r * stride + (uint64_t)gchunk * 2 : base;
        tile::cp_async16_zfill(smem_stage + (uint32_t)row * SLAB_BYTES + (uint32_t)((ch ^ (row & 7)) << 4), src, bytes);
    }
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// 32-bit word w (0 .. 31) of staged row r
__device__ __forceinline__ uint32_t word_addr(uint32_t stage, uint32_t r, uint32_t w) {
    return stage + r * SLAB_BYTES + ((((w >> 2) ^ (r & 7u)) << 4) | ((w & 3u) << 2));
}

__global__ void __launch_bounds__(THREADS, 1) dense_b1_kernel(const DenseJob job) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ unsigned long long warp_part[THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t g = lane >> 2, t = lane & 3;             // fragment coordinates (PTX ISA, m16n8k256 .b1)
    const uint32_t wm = warp & 3, wn = warp >> 2;           // warp tile: A rows [32 wm, +32), B rows [64 wn, +64)

    uint32_t bi, bj;
    tile_coords(job, job.tile_begin + blockIdx.x, TM, TN, bi, bj);
    const uint64_t rowA0 = (uint64_t)bi * TM, rowB0 = (uint64_t)bj * TN;
    const uint32_t n_slabs = (job.n_words * 8 + SLAB_BYTES - 1) / SLAB_BYTES;
    const uint32_t smem_base = smem_u32(smem);

    uint32_t acc[2][8][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[m][n][k] = 0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if ((uint32_t)s < n_slabs) load_slab(job, smem_base + s * STAGE_BYTES, rowA0, rowB0, s);
        tile::cp_async_commit();
    }
    for (uint32_t slab = 0; slab < n_slabs; ++slab) {
        tile::cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const uint32_t nxt = slab + STAGES - 1;
            if (nxt < n_slabs) load_slab(job, smem_base + (nxt % STAGES) * STAGE_BYTES, rowA0, rowB0, nxt);
            tile::cp_async_commit();
        }
        const uint32_t stage = smem_base + (slab % STAGES) * STAGE_BYTES;
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {               // k256 steps of the slab: words [8 ks, 8 ks + 8)
            uint32_t a[2][4];
#pragma unroll
            for (uint32_t m = 0; m < 2; ++m) {
                const uint32_t r = wm * 32 + m * 16 + g;
                a[m][0] = lds32(word_addr(stage, r, ks * 8 + t));
                a[m][1] = lds32(word_addr(stage, r + 8, ks * 8 + t));
                a[m][2] = lds32(word_addr(stage, r, ks * 8 + 4 + t));
                a[m][3] = lds32(word_addr(stage, r + 8, ks * 8 + 4 + t));
            }
#pragma unroll
            for (uint32_t n = 0; n < 8; ++n) {
                const uint32_t r = TM + wn * 64 + n * 8 + g;
                const uint32_t b0 = lds32(word_addr(stage, r, ks * 8 + t)), b1v = lds32(word_addr(stage, r, ks * 8 + 4 + t));
                mma_b1(acc[0][n], a[0], b0, b1v);
                mma_b1(acc[1][n], a[1], b0, b1v);
            }
        }
    }
    tile::cp_async_wait<0>();

    // ---- epilogue: c0/c1 = (row g, cols 2t, 2t + 1), c2/c3 = (row g + 8, same cols) ----------------
    unsigned long long sum = 0;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t li = rowA0 + wm * 32 + m * 16 + g + (k >> 1) * 8;
                const uint64_t lj = rowB0 + wn * 64 + n * 8 + 2 * t + (k & 1);
                if (li >= job.nA || lj >= job.nB) continue;
                uint32_t v = acc[m][n][k];
                if (job.strict_upper && job.j_off + lj <= job.i_off + li) v = 0;
                sum += v;
                if (job.out) job.out[li * job.ld + lj] = v;
            }
    if (job.total) {
        sum = warp_sum(sum);
        if (lane == 0) warp_part[warp] = sum;
        __syncthreads();
        if (tid == 0) {
            unsigned long long tot = 0;
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) tot += warp_part[w];
            if (tot) atomicAdd(job.total, tot);
        }
    }
}

}  // namespace b1

TileShape b1_tile_shape() { return {b1::TM, b1::TN}; }

int launch_dense_b1(const DenseJob& job, cudaStream_t stream) {
    if (job.tile_end <= job.tile_begin) return STORM_B200_OK;
    const int smem_bytes = b1::STAGES * b1::STAGE_BYTES;
    STORM_CUDA_TRY(cudaFuncSetAttribute(b1::dense_b1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    uint64_t remaining = job.tile_end - job.tile_begin, begin = job.tile_begin;
    while (remaining) {                                // grid.x is limited to 2^31 - 1
        const uint64_t n = remaining > 0x40000000ull ? 0x40000000ull : remaining;
        DenseJob j = job;
        j.tile_begin = begin;
        j.tile_end = begin + n;
        b1::dense_b1_kernel<<<(unsigned)n, b1::THREADS, smem_bytes, stream>>>(j);
        STORM_CUDA_TRY(cudaGetLastError());
        count_launch();
        begin += n;
        remaining -= n;
    }
    return STORM_B200_OK;
}

}  // namespace storm
