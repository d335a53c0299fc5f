// dense_umma.cu -- tensor-core tile kernel: tcgen05.mma kind::i8 over bits that are
// unpacked to {0,1} bytes on the fly.
//
// XX^T over a binary matrix is an integer GEMM: popcount(row_i & row_j) =
// sum_k bit(i,k) * bit(j,k).  Replaces the same loop nest as dense_popc.cu
// (storm.c:1165-1169 / 1199-1238 + libalgebra.h:2684-2744) with one UMMA tile per
// CTA (pair).  Accumulation is exact: u8 x u8 products into s32 accumulators in
// tensor memory, and a pair count is at most M < 2^31.
//
// Structure, per CTA:
//   warps 0-3   "A expanders": thread = one A row = one TMEM lane.  Each k-block
//               (128 bits) is loaded packed (16 B, ld.global.nc), expanded with
//               (w >> j) & 0x01010101 into 32 registers and written to tensor memory
//               with tcgen05.st (A operand lives in TMEM: no shared-memory traffic).
//               After the K loop the same warps run the epilogue (tcgen05.ld).
//   warps 4..   "B expanders": thread = one B row.  Same expansion, written with
//               st.shared.v4 into the canonical K-major SWIZZLE_128B layout the UMMA
//               shared-memory descriptor expects (16-byte chunk c of row r lands at
//               chunk c ^ (r & 7) of its 128-byte line; 8-row groups are 1024 B apart).
//   last warp   TMEM allocation and, in the leader CTA, the single thread that issues
//               tcgen05.mma (4 per k-block, K = 32 bytes each) and tcgen05.commit.
//   Stages are handed over with mbarriers: full[s] (expanders -> MMA), empty[s]
//   (tcgen05.commit -> expanders), acc_full (last commit -> epilogue).
//
// Bit order: within a 32-bit word, output register j holds bits j, j+8, j+16, j+24 as
// its four bytes.  A and B use the same permutation of K, and a dot product is
// invariant under a common permutation of its terms, so no un-shuffling is needed.
//
// CG = 2 (cta_group::2): two CTAs of a cluster share one 256 x 256 tile; each holds
// 128 A rows in its own TMEM and supplies 128 of the 256 B rows from its own shared
// memory, which halves the per-SM expansion work and shared-memory reads per MMA.
#include "common.cuh"

namespace storm {
namespace {

constexpr int UM_N = 256;               // B rows (accumulator columns) per tile
constexpr int UM_TMEM_COLS = 512;
constexpr int UM_ACC_COL = 0;           // accumulator: columns [0, 256)
constexpr int UM_A_COL = 256;           // A stage s: columns [256 + 32 s, 256 + 32 s + 32)
constexpr int UM_PREFETCH = 4;          // packed k-blocks each expander keeps in registers

template <int CG>
struct Cfg {
    static constexpr int A_WARPS = 4;
    // k-blocks in flight.  The pair kernel hands stages over through cluster-scope barriers (remote
    // arrive, multicast commit), so it needs a deeper ring to cover that latency; 8 stages of A fill
    // the 256 TMEM columns next to the accumulator.
    static constexpr int STAGES = CG == 2 ? 8 : 6;
    static constexpr int B_ROWS = UM_N / CG;                 // B rows expanded by this CTA
    static constexpr int B_WARPS = B_ROWS / 32;
    static constexpr int MMA_WARP = A_WARPS + B_WARPS;
    static constexpr int THREADS = (A_WARPS + B_WARPS + 1) * 32;
    static constexpr int STAGE_BYTES = B_ROWS * 128;         // expanded B rows of one k-block
    static constexpr int TM = 128 * CG, TN = UM_N;
    static constexpr int PRODUCER_ARRIVALS = CG * (A_WARPS + B_WARPS);
    static constexpr uint32_t SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + 256;
    // kind::i8 instruction descriptor (cute::UMMA::InstrDescriptor bit layout):
    //   [4,6) c_format = 2 (S32); [7,10) a_format = 0 (u8); [10,13) b_format = 0 (u8);
    //   [15] a_major = 0 (K); [16] b_major = 0 (K); [17,23) N >> 3; [24,29) M >> 4
    static constexpr uint32_t IDESC = (2u << 4) | ((uint32_t)(UM_N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
};

// ---- PTX wrappers -----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Bounded spin: a protocol bug traps (CUDA error) instead of hanging the device.
// Default (.acquire.cta) semantics on purpose: an explicit .acquire.cluster makes ptxas emit
// CCTL.IVALL (L1 invalidate) per wait and .release.cluster a MEMBAR.ALL.GPU per arrive, which
// more than halved the pair kernel.  The data handed over is ordered by its own fences
// (tcgen05.wait::st + tcgen05.fence for TMEM, fence.proxy.async for shared memory).
template <bool CLUSTER>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 r;\n\t"
                 "mapa.shared::cluster.u32 r, %0, %1;\n\t"
                 "mbarrier.arrive.shared::cluster.b64 _, [r];\n\t}"
                 ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
    if (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_free(uint32_t taddr) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(UM_TMEM_COLS) : "memory");
    else         asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(UM_TMEM_COLS) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T, A and B K-major u8, D s32.
template <int CG>
__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on `bar` (in every CTA of the pair for CG = 2) once all MMAs issued so far have completed.
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// One packed k-block (128 bits) of a row; `words` = how many of its 64-bit words carry data.
__device__ __forceinline__ uint4 load_kblock(const uint4* row, uint32_t kb, uint32_t n_kb, uint32_t n_words) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row == nullptr || kb >= n_kb) return v;
    const uint4* p = row + kb;
    if (2 * kb + 2 <= n_words) {
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    } else {                                   // last k-block of an odd-width row: one word only
        asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    }
    return v;
}

// 32 bits -> 32 bytes of {0,1}: register j holds bits j, j+8, j+16, j+24.
__device__ __forceinline__ void expand32(uint32_t w, uint32_t (&r)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = (w >> j) & 0x01010101u;
}

template <int CG>
__global__ void __launch_bounds__(Cfg<CG>::THREADS, 1) dense_umma_kernel(const DenseJob job) {
    using C = Cfg<CG>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B needs 1024-byte alignment
    const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
    const uint32_t full_bar = bar_base;                                     // STAGES x 8 B
    const uint32_t empty_bar = bar_base + 8 * C::STAGES;
    const uint32_t acc_bar = bar_base + 16 * C::STAGES;
    const uint32_t tmem_slot = acc_bar + 8;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));        // generic pointer to the aligned base
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + C::STAGES * C::STAGE_BYTES + 16 * C::STAGES + 8);
    unsigned long long* red = reinterpret_cast<unsigned long long*>(smem_gen + C::STAGES * C::STAGE_BYTES + 16 * C::STAGES + 16);

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const uint64_t tile = job.tile_begin + (CG == 2 ? (blockIdx.x >> 1) : blockIdx.x);
    uint32_t bi, bj;
    tile_coords(job, tile, C::TM, C::TN, bi, bj);
    const uint64_t rowA0 = (uint64_t)bi * C::TM + rank * 128u;              // this CTA's 128 A rows
    const uint64_t rowB0 = (uint64_t)bj * C::TN;                            // the tile's 256 B rows
    const uint32_t n_kb = (job.n_words + 1) / 2;

    // ---- setup ----------------------------------------------------------------
    if (warp == C::MMA_WARP) tmem_alloc<CG>(tmem_slot);
    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar + 8 * s, C::PRODUCER_ARRIVALS);
            mbar_init(empty_bar + 8 * s, 1);
        }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < C::A_WARPS) {
        // ===== A expander: row -> TMEM lane ========================================
        const uint64_t r = rowA0 + warp * 32 + lane;
        const uint4* src = (r < job.nA) ? reinterpret_cast<const uint4*>(job.A + r * job.strideA) : nullptr;
        const uint32_t lane_base = tmem_base + ((warp * 32u) << 16);
        uint4 pf[UM_PREFETCH];
#pragma unroll
        for (int u = 0; u < UM_PREFETCH; ++u) pf[u] = load_kblock(src, u, n_kb, job.n_words);
        for (uint32_t kb0 = 0; kb0 < n_kb; kb0 += UM_PREFETCH) {
#pragma unroll
            for (int u = 0; u < UM_PREFETCH; ++u) {
                const uint32_t kb = kb0 + u;
                if (kb < n_kb) {
                    const uint32_t s = kb % C::STAGES, it = kb / C::STAGES;
                    const uint4 w = pf[u];
                    pf[u] = load_kblock(src, kb + UM_PREFETCH, n_kb, job.n_words);
                    mbar_wait<CG == 2>(empty_bar + 8 * s, (it & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t t = lane_base + UM_A_COL + s * 32;
                    uint32_t e[8];
                    expand32(w.x, e); tmem_st8(t + 0, e);
                    expand32(w.y, e); tmem_st8(t + 8, e);
                    expand32(w.z, e); tmem_st8(t + 16, e);
                    expand32(w.w, e); tmem_st8(t + 24, e);
                    tc_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 2) mbar_arrive_cluster(full_bar + 8 * s, 0); else mbar_arrive_local(full_bar + 8 * s);
                    }
                }
            }
        }
    } else if (warp < C::MMA_WARP) {
        // ===== B expander: row -> swizzled shared-memory line =======================
        const uint32_t idx = tid - C::A_WARPS * 32;                        // 0 .. B_ROWS-1
        const uint64_t r = rowB0 + rank * C::B_ROWS + idx;
        const uint4* src = (r < job.nB) ? reinterpret_cast<const uint4*>(job.B + r * job.strideB) : nullptr;
        const uint32_t line = (idx >> 3) * 1024u + (idx & 7u) * 128u;      // 8-row groups are 1024 B apart
        const uint32_t sw = idx & 7u;
        uint4 pf[UM_PREFETCH];
#pragma unroll
        for (int u = 0; u < UM_PREFETCH; ++u) pf[u] = load_kblock(src, u, n_kb, job.n_words);
        for (uint32_t kb0 = 0; kb0 < n_kb; kb0 += UM_PREFETCH) {
#pragma unroll
            for (int u = 0; u < UM_PREFETCH; ++u) {
                const uint32_t kb = kb0 + u;
                if (kb < n_kb) {
                    const uint32_t s = kb % C::STAGES, it = kb / C::STAGES;
                    const uint4 w = pf[u];
                    pf[u] = load_kblock(src, kb + UM_PREFETCH, n_kb, job.n_words);
                    mbar_wait<CG == 2>(empty_bar + 8 * s, (it & 1) ^ 1);
                    const uint32_t dst = smem_base + s * C::STAGE_BYTES + line;
                    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                          // K step k = bytes [32k, 32k+32) of the line
                        uint32_t e[8];
                        expand32(ws[k], e);
                        st_shared_v4(dst + (((2 * k) ^ sw) << 4), e[0], e[1], e[2], e[3]);
                        st_shared_v4(dst + (((2 * k + 1) ^ sw) << 4), e[4], e[5], e[6], e[7]);
                    }
                    fence_proxy_async_smem();                              // generic writes -> visible to the UMMA (async proxy)
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 2) mbar_arrive_cluster(full_bar + 8 * s, 0); else mbar_arrive_local(full_bar + 8 * s);
                    }
                }
            }
        }
    } else if (rank == 0 && lane == 0) {
        // ===== MMA issuer: one thread of the leader CTA ==============================
        // K-major SWIZZLE_128B shared-memory descriptor (cute::UMMA::SmemDescriptor):
        //   [0,14) addr >> 4; [16,30) LBO >> 4 = 1 (unused for swizzled K-major); [32,46) SBO >> 4 = 64 (1024 B
        //   between 8-row groups); [46,48) version = 1; [61,64) layout = 2 (SWIZZLE_128B)
        const uint64_t desc_hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        for (uint32_t kb = 0; kb < n_kb; ++kb) {
            const uint32_t s = kb % C::STAGES, it = kb / C::STAGES;
            mbar_wait<CG == 2>(full_bar + 8 * s, it & 1);
            tc_fence_after();
            const uint32_t b_addr = smem_base + s * C::STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t b_desc = desc_hi | (uint64_t)(((b_addr + k * 32) >> 4) & 0x3FFF);
                umma_i8_ts<CG>(tmem_base + UM_ACC_COL, tmem_base + UM_A_COL + s * 32 + k * 8, b_desc, C::IDESC,
                               (kb | (uint32_t)k) != 0);
            }
            umma_commit<CG>(empty_bar + 8 * s);                            // frees the stage when these MMAs are done
        }
        umma_commit<CG>(acc_bar);                                          // accumulator complete
    }
    __syncwarp();                                                          // re-converge the MMA warp (aligned ops follow)

    // ---- epilogue: TMEM -> registers -> masked sum / per-pair store ---------------
    unsigned long long sum = 0;
    if (warp < C::A_WARPS) {
        mbar_wait<CG == 2>(acc_bar, 0);
        tc_fence_after();
        const uint64_t li = rowA0 + warp * 32 + lane;                      // A row of this thread (= its TMEM lane)
        const uint64_t gi = job.i_off + li;
        const bool row_ok = li < job.nA;
        const uint32_t lane_base = tmem_base + ((warp * 32u) << 16) + UM_ACC_COL;
#pragma unroll 1
        for (int c0 = 0; c0 < UM_N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(lane_base + c0, v);
            tc_wait_ld();
            if (row_ok) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const uint64_t lj = rowB0 + c0 + c;
                    if (lj < job.nB) {
                        uint32_t x = v[c];
                        if (job.strict_upper && job.j_off + lj <= gi) x = 0;
                        sum += x;
                        if (job.out) job.out[li * job.ld + lj] = x;
                    }
                }
            }
        }
    }
    if (job.total) {
        sum = warp_sum(sum);
        if (lane == 0 && warp < C::A_WARPS) red[warp] = sum;
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();                 // everyone is done with TMEM / smem
    if (job.total && tid == 0) {
        const unsigned long long t = red[0] + red[1] + red[2] + red[3];
        if (t) atomicAdd(job.total, t);
    }
    if (warp == C::MMA_WARP) tmem_free<CG>(tmem_base);
}

template <int CG>
int launch_cg(const DenseJob& job, cudaStream_t stream) {
    using C = Cfg<CG>;
    STORM_CUDA_TRY(cudaFuncSetAttribute(dense_umma_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    uint64_t remaining = job.tile_end - job.tile_begin, begin = job.tile_begin;
    while (remaining) {
        const uint64_t n = remaining > 0x20000000ull ? 0x20000000ull : remaining;
        DenseJob j = job;
        j.tile_begin = begin;
        j.tile_end = begin + n;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(n * CG));
        cfg.blockDim = dim3(C::THREADS);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CG;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        STORM_CUDA_TRY(cudaLaunchKernelEx(&cfg, dense_umma_kernel<CG>, j));
        count_launch();
        begin += n;
        remaining -= n;
    }
    return STORM_B200_OK;
}

int g_umma_cg = 2;   // cta_group used by launch_dense_umma (1 or 2); see STORM_b200_set_umma_cta_group

}  // namespace

TileShape umma_tile_shape() { return {(uint32_t)(128 * g_umma_cg), (uint32_t)UM_N}; }

bool umma_supports(const DenseJob& job) {
    if (job.n_words == 0 || job.n_words >= (1u << 25)) return false;        // counts stay below 2^31
    if ((job.strideA & 1) || (job.strideB & 1)) return false;
    if (job.A && ((uintptr_t)job.A & 15)) return false;
    if (job.B && ((uintptr_t)job.B & 15)) return false;
    return true;
}

int launch_dense_umma(const DenseJob& job, cudaStream_t stream) {
    if (job.tile_end <= job.tile_begin) return STORM_B200_OK;
    return g_umma_cg == 2 ? launch_cg<2>(job, stream) : launch_cg<1>(job, stream);
}

}  // namespace storm

// Development / measurement knob: cta_group of the UMMA kernel (1 = one CTA per 128 x 256 tile,
// 2 = CTA pair per 256 x 256 tile).  Returns the previous value.
extern "C" int STORM_b200_set_umma_cta_group(int cg) {
    const int prev = storm::g_umma_cg;
    if (cg == 1 || cg == 2) storm::g_umma_cg = cg;
    return prev;
}
