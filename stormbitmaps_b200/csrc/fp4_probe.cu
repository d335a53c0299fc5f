// fp4_probe.cu -- is tcgen05.mma kind::mxf4 usable for EXACT bit counting, and how fast is it?
//
// The tensor formulation of the path (dense_umma.cu) multiplies bits that were unpacked to bytes
// with kind::i8.  Blackwell's block-scaled FP4 kind runs at twice that rate (K = 64 four-bit
// elements per instruction instead of 32 bytes) and E2M1 holds 0, 0.5, 1 and 2 exactly, so bits
// unpacked to NIBBLES would halve both the expansion work and the tensor time per bit -- if, and
// only if, the fp32 accumulation inside the tensor core is exact for integer sums up to M.
// That is a property of the hardware, so it is measured, not assumed:
//
//   fp4_exact_kernel   one CTA, cta_group::1, M 128 x N 256 x K 64, A from tensor memory, B from
//                      SWIZZLE_128B shared memory, all scale factors UE8M0 = 127 (x 1.0).  Runs a
//                      list of (n_full, n_single, pattern) cases: n_full instructions that add 64 to
//                      every accumulator followed by n_single that add exactly 1, then compares
//                      all 128 x 256 accumulators with 64 n_full + n_single.  Patterns cover the
//                      operand encodings the tile kernel would use (1.0 x 1.0, 0.5 x 2.0, 2.0 x 0.5)
//                      and the nibble <-> K-index correspondence between TMEM A and shared B.
//   fp4_peak_kernel    the same instruction back to back on every SM: the pipe's ceiling.
//
// Results go to STORM_b200_fp4_probe() / STORM_b200_microbench(6|7); tools/fp4_probe.py prints them.
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "runtime.h"
#include "umma_ptx.cuh"

namespace storm {
namespace {

constexpr int FP_N = 256;
constexpr int FP_ACC_COL = 0;
constexpr int FP_A_COL = 256;            // operand patterns: 8 columns each
constexpr int FP_SF_COL = 384;           // 64 columns of 0x7F7F7F7F (every layout of SFA / SFB reads 1.0)

// Block-scaled instruction descriptor (cute::UMMA::InstrDescriptorBlockScaled): [7,10) a_format = 1
// (E2M1), [10,13) b_format = 1, [15]/[16] K-major, [17,23) N >> 3, [23] scale format 1 = UE8M0,
// [24,29) M >> 4, [31] k_size 0 = K 64.
template <int CG, int N = FP_N>
__host__ __device__ constexpr uint32_t fp4_idesc() {
    return (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)((128 * CG) >> 4) << 24);
}

template <int CG>
__device__ __forceinline__ void umma_mxf4_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t sfa, uint32_t sfb, uint32_t accumulate) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%5], [%6], p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%5], [%6], p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb) : "memory");
}

struct Fp4Case {
    uint32_t n_full, n_single, pattern;
};
struct Fp4Result {
    float expect, got_min, got_max;
    uint32_t mismatches;
};

// pattern -> (A nibble word, B byte) for the "full" instructions
//   0: 1.0 x 1.0   1: 0.5 x 2.0   2: 2.0 x 0.5   3: alternating (0.5, 2.0) x (2.0, 0.5)
__device__ __forceinline__ uint32_t pat_a(uint32_t p) { return p == 0 ? 0x22222222u : p == 1 ? 0x11111111u : p == 2 ? 0x44444444u : 0x41414141u; }
__device__ __forceinline__ uint32_t pat_b(uint32_t p) { return p == 0 ? 0x22222222u : p == 1 ? 0x44444444u : p == 2 ? 0x11111111u : 0x14141414u; }

__global__ void __launch_bounds__(128, 1) fp4_exact_kernel(const Fp4Case* cases, Fp4Result* results, uint32_t n_cases) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    constexpr uint32_t B_BYTES = FP_N * 128;                               // 256 rows x one 128-byte swizzle line
    const uint32_t b_full = smem_base, b_single = smem_base + B_BYTES;
    const uint32_t bar = smem_base + 2 * B_BYTES;
    const uint32_t tmem_slot = bar + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    uint32_t* red = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot + 8 - smem_base));   // [4][3]
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) tmem_alloc<1>(tmem_slot);
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t my_lanes = tmem_base + ((warp * 32u) << 16);

    // scale factors: x 1.0 everywhere
    {
        uint32_t sf[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sf[j] = 0x7F7F7F7Fu;
        for (int c = 0; c < 64; c += 8) tmem_st8(my_lanes + FP_SF_COL + c, sf);
        tc_wait_st();
    }
    const uint64_t desc_hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
    uint32_t phase = 0;

    for (uint32_t ci = 0; ci < n_cases; ++ci) {
        const Fp4Case cs = cases[ci];
        // A "full": 8 columns of the pattern word.  A "single": K index 19 only (column 2, nibble 3) = 1.0.
        {
            uint32_t a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = pat_a(cs.pattern);
            tmem_st8(my_lanes + FP_A_COL, a);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = 0;
            a[2] = 0x2u << 12;
            tmem_st8(my_lanes + FP_A_COL + 8, a);
            tc_wait_st();
        }
        // B "full": every byte the pattern.  B "single": K index 19 only = byte 9 high nibble = 1.0, i.e.
        // byte 9 of the row's first 32-byte K step: logical 16-byte chunk 0, swizzled with the row.
        for (uint32_t i = tid; i < B_BYTES / 16; i += 128) {
            const uint32_t pb = pat_b(cs.pattern);
            st_shared_v4(b_full + i * 16, pb, pb, pb, pb);
            st_shared_v4(b_single + i * 16, 0, 0, 0, 0);
        }
        __syncthreads();
        for (uint32_t r = tid; r < (uint32_t)FP_N; r += 128) {
            const uint32_t line = b_single + (r >> 3) * 1024u + (r & 7u) * 128u;
            const uint32_t chunk = (0u ^ (r & 7u)) << 4;
            st_shared_v4(line + chunk, 0, 0, 0x20u << 8, 0);               // byte 9 = 0x20: high nibble (K 19) = 1.0
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint64_t bd_full = desc_hi | (uint64_t)((b_full >> 4) & 0x3FFF);
            const uint64_t bd_single = desc_hi | (uint64_t)((b_single >> 4) & 0x3FFF);
            const uint32_t sfa = tmem_base + FP_SF_COL, sfb = tmem_base + FP_SF_COL + 32;
            uint32_t first = 0;
            for (uint32_t i = 0; i < cs.n_full; ++i, first = 1)
                umma_mxf4_ts<1>(tmem_base + FP_ACC_COL, tmem_base + FP_A_COL, bd_full, fp4_idesc<1>(), sfa, sfb, first);
            for (uint32_t i = 0; i < cs.n_single; ++i, first = 1)
                umma_mxf4_ts<1>(tmem_base + FP_ACC_COL, tmem_base + FP_A_COL + 8, bd_single, fp4_idesc<1>(), sfa, sfb, first);
            umma_commit<1>(bar);
        }
        mbar_wait_t<true>(bar, phase);
        phase ^= 1;
        tc_fence_after();
        const float expect = 64.0f * (float)cs.n_full + (float)cs.n_single;
        float mn = 3.0e38f, mx = -3.0e38f;
        uint32_t bad = 0;
        for (uint32_t c0 = 0; c0 < (uint32_t)FP_N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(my_lanes + FP_ACC_COL + c0, v);
            tc_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float f = __uint_as_float(v[k]);
                mn = fminf(mn, f); mx = fmaxf(mx, f);
                bad += (f != expect);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            bad += __shfl_xor_sync(0xffffffffu, bad, o);
        }
        if (lane == 0) { red[warp * 3] = __float_as_uint(mn); red[warp * 3 + 1] = __float_as_uint(mx); red[warp * 3 + 2] = bad; }
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            Fp4Result r{expect, 3.0e38f, -3.0e38f, 0};
            for (int w = 0; w < 4; ++w) {
                r.got_min = fminf(r.got_min, __uint_as_float(red[w * 3]));
                r.got_max = fmaxf(r.got_max, __uint_as_float(red[w * 3 + 1]));
                r.mismatches += red[w * 3 + 2];
            }
            results[ci] = r;
        }
        __syncthreads();
        tc_fence_after();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free<1>(tmem_base);
}

template <int CG, int N>
__global__ void __launch_bounds__(128, 1) fp4_peak_kernel(uint32_t iters, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    constexpr uint32_t RING = 8;                                             // commits in flight
    constexpr uint32_t BATCH = 8;                                            // MMAs per commit
    constexpr uint32_t B_BYTES = (N / CG) * 128;
    const uint32_t bar_base = smem_base + B_BYTES;
    const uint32_t tmem_slot = bar_base + 8 * RING;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;

    for (uint32_t i = tid; i < B_BYTES / 16; i += blockDim.x) st_shared_v4(smem_base + i * 16, 0, 0, 0, 0);
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc<CG>(tmem_slot);
    if (tid == 0) {
        for (uint32_t b = 0; b < RING; ++b) mbar_init(bar_base + 8 * b, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // warp-uniform issue loop with one elected lane, as in the tile kernel (see umma_peak_kernel)
    if (rank == 0 && warp == 0) {
        const bool leader = elect_one();
        const uint64_t desc_hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t sfa = tmem_u + FP_SF_COL, sfb = tmem_u + FP_SF_COL + 32;
        const long long c0 = clock64();
        for (uint32_t it = 0; it < iters; ++it) {
            if (it >= RING) mbar_wait_t<true>(bar_base + 8 * (it % RING), ((it / RING) - 1) & 1);
            if (leader) {
#pragma unroll
                for (int k = 0; k < (int)BATCH; ++k) {
                    const uint64_t b_desc = desc_hi | (uint64_t)(((smem_base + (k & 3) * 32) >> 4) & 0x3FFF);
                    umma_mxf4_ts<CG>(tmem_u + FP_ACC_COL, tmem_u + FP_A_COL + (k & 3) * 8, b_desc, fp4_idesc<CG, N>(), sfa, sfb, 1u);
                }
                umma_commit<CG>(bar_base + 8 * (it % RING));
            }
            __syncwarp();
        }
        for (uint32_t it = iters > RING ? iters - RING : 0; it < iters; ++it)
            mbar_wait_t<true>(bar_base + 8 * (it % RING), (it / RING) & 1);
        const long long c1 = clock64();
        if (leader && cycles) cycles[blockIdx.x / CG] = (unsigned long long)(c1 - c0);
    }
    __syncwarp();
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 0) tmem_free<CG>(tmem_base);
}

template <int CG, int N>
int run_fp4_peak(double* ops_per_s, double* clock64_mhz) {
    int dev = 0, sms = 0;
    STORM_CUDA_TRY(cudaGetDevice(&dev));
    STORM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int smem_bytes = 1024 + (N / CG) * 128 + 256;
    STORM_CUDA_TRY(cudaFuncSetAttribute(fp4_peak_kernel<CG, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(sms / CG * CG));
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    STORM_CUDA_TRY(cudaEventCreate(&e0));
    STORM_CUDA_TRY(cudaEventCreate(&e1));
    const unsigned n_clusters = cfg.gridDim.x / CG;
    unsigned long long* d_cyc = nullptr;
    STORM_CUDA_TRY(cudaMalloc(&d_cyc, n_clusters * sizeof(unsigned long long)));
    std::vector<unsigned long long> h_cyc(n_clusters);
    const uint32_t iters = 50000;
    double best = 0, best_mhz = 0;
    for (int rep = 0; rep < 4; ++rep) {                                   // rep 0 is the warm-up
        STORM_CUDA_TRY(cudaEventRecord(e0));
        STORM_CUDA_TRY(cudaLaunchKernelEx(&cfg, fp4_peak_kernel<CG, N>, iters, d_cyc));
        STORM_CUDA_TRY(cudaEventRecord(e1));
        STORM_CUDA_TRY(cudaEventSynchronize(e1));
        count_launch();
        float ms = 0;
        STORM_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        STORM_CUDA_TRY(cudaMemcpy(h_cyc.data(), d_cyc, n_clusters * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        double cyc = 0;
        for (unsigned long long c : h_cyc) cyc += (double)c;
        cyc /= n_clusters;
        // per SM and instruction: 128 x N x 64 MACs = 2 ops each
        const double ops = (double)cfg.gridDim.x * iters * 8.0 * 128.0 * (double)N * 64.0 * 2.0;
        if (rep > 0 && ops / (ms * 1e-3) > best) { best = ops / (ms * 1e-3); best_mhz = cyc / (ms * 1e-3) / 1e6; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_cyc);
    *ops_per_s = best;
    if (clock64_mhz) *clock64_mhz = best_mhz;
    return STORM_B200_OK;
}

// ---- data-dependent increments, cta_group 1 and 2 ----------------------------------------------------------
// The probe above drives every accumulator with the SAME increment per instruction (+64, then +1) on one CTA.
// The tile kernel is cta_group::2 and its increments are whatever the data says: 0 .. 64 per instruction,
// different for every accumulator element.  This kernel reproduces that: 4 A operands (tensor memory) x 4 B
// operands (the four K steps of one SWIZZLE_128B line block) hold pseudo-random bit masks in the production
// encoding (nibble codes {0.5, 1, 2, 1} x {2, 1, 0.5, 1} by register of a word, expand32_*_fp4 of dense_umma.cu),
// with all-ones and all-zero rows mixed in; a pseudo-random sequence of n_steps (A, B) combinations is issued
// back to back into one accumulator, and every element is compared with the integer it must hold:
//     sum over (a, b) of times[a][b] x popcount(maskA[a][row] & maskB[b][col]).
// Row 0 x column 0 are all ones in every operand, so that element reaches 64 x n_steps (2^24 - 64 at 262 143 steps).
struct Fp4RandResult {
    float max_expected, min_diff, max_diff;
    uint32_t mismatches;
};

__host__ __device__ inline uint64_t probe_hash(uint64_t x) {               // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// 64-bit mask of operand `op` (0..3) of side (0 = A, 1 = B) for global row / column idx
__host__ __device__ inline uint64_t probe_mask(uint32_t seed, uint32_t side, uint32_t op, uint32_t idx) {
    const uint32_t cls = idx & 15u;
    if (cls == 0) return ~0ull;                                            // every instruction adds 64 where two of these meet
    if (cls == 1) return 0ull;
    const uint64_t h1 = probe_hash(((uint64_t)seed << 32) ^ ((uint64_t)side << 28) ^ ((uint64_t)op << 24) ^ idx);
    const uint64_t h2 = probe_hash(h1 ^ 0xD1B54A32D192ED03ull);
    return cls < 6 ? (h1 & h2) : cls < 11 ? h1 : (h1 | h2);                 // ~25 %, 50 %, 75 % ones
}
// 32 mask bits -> 4 TMEM cells / shared-memory words of 8 E2M1 nibbles each, production encoding
__device__ __forceinline__ void probe_encode(uint32_t bits, bool b_side, uint32_t (&cell)[4]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t w = 0;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const uint32_t k = 8u * c + n;                                 // K index within this half
            // as the tile kernel lays a word out: cell (register) c of the four holds one code in all its nibbles
            const uint32_t code = b_side ? (c == 0 ? 4u : c == 2 ? 1u : 2u)                  // 2, 1, 0.5, 1
                                         : (c == 0 ? 1u : c == 2 ? 4u : 2u);                 // 0.5, 1, 2, 1
            if ((bits >> k) & 1u) w |= code << (4 * n);
        }
        cell[c] = w;
    }
}
__host__ __device__ inline uint32_t probe_step_combo(uint32_t seed, uint32_t t) { return (uint32_t)(probe_hash(((uint64_t)seed << 32) | t) >> 17) & 15u; }

template <int CG>
__global__ void __launch_bounds__(128, 1) fp4_random_kernel(uint32_t n_steps, uint32_t seed, Fp4RandResult* results) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    constexpr uint32_t B_ROWS = FP_N / CG;                                 // B rows (accumulator columns) held by this CTA
    constexpr uint32_t B_BYTES = B_ROWS * 128;
    const uint32_t bar = smem_base + B_BYTES;
    const uint32_t tmem_slot = bar + 8;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    uint32_t* times = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot + 8 - smem_base));     // [16]
    uint32_t* red = times + 16;                                                                  // [4][4]
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;

    if (warp == 0) tmem_alloc<CG>(tmem_slot);
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (tid < 16) times[tid] = 0;
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t my_lanes = tmem_base + ((warp * 32u) << 16);
    const uint32_t row = rank * 128u + tid;                                // global accumulator row of this thread (= its TMEM lane)

    {   // scale factors x 1.0; A operands: operand a occupies columns [FP_A_COL + 8 a, + 8)
        uint32_t sf[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sf[j] = 0x7F7F7F7Fu;
        for (int c = 0; c < 64; c += 8) tmem_st8(my_lanes + FP_SF_COL + c, sf);
        for (uint32_t a = 0; a < 4; ++a) {
            const uint64_t m = probe_mask(seed, 0, a, row);
            uint32_t lo[4], hi[4], cells[8];
            probe_encode((uint32_t)m, false, lo);
            probe_encode((uint32_t)(m >> 32), false, hi);
#pragma unroll
            for (int c = 0; c < 4; ++c) { cells[c] = lo[c]; cells[4 + c] = hi[c]; }
            tmem_st8(my_lanes + FP_A_COL + 8 * a, cells);
        }
        tc_wait_st();
    }
    // B operands: K step b (32 bytes) of the 128-byte line of B row r holds operand b of global column rank * B_ROWS + r
    for (uint32_t r = tid; r < B_ROWS; r += 128) {
        const uint32_t line = smem_base + (r >> 3) * 1024u + (r & 7u) * 128u;
        for (uint32_t b = 0; b < 4; ++b) {
            const uint64_t m = probe_mask(seed, 1, b, rank * B_ROWS + r);
            uint32_t lo[4], hi[4];
            probe_encode((uint32_t)m, true, lo);
            probe_encode((uint32_t)(m >> 32), true, hi);
            st_shared_v4(line + (((2 * b) ^ (r & 7u)) << 4), lo[0], lo[1], lo[2], lo[3]);
            st_shared_v4(line + (((2 * b + 1) ^ (r & 7u)) << 4), hi[0], hi[1], hi[2], hi[3]);
        }
    }
    fence_proxy_async_smem();
    if (tid == 0) for (uint32_t t = 0; t < n_steps; ++t) times[probe_step_combo(seed, t)] += 1;   // both CTAs: the same sequence
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();

    if (rank == 0 && tid == 0) {
        const uint64_t desc_hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t sfa = tmem_base + FP_SF_COL, sfb = tmem_base + FP_SF_COL + 32;
        for (uint32_t t = 0; t < n_steps; ++t) {
            const uint32_t combo = probe_step_combo(seed, t), a = combo >> 2, b = combo & 3u;
            const uint64_t b_desc = desc_hi | (uint64_t)(((smem_base + b * 32) >> 4) & 0x3FFF);
            umma_mxf4_ts<CG>(tmem_base + FP_ACC_COL, tmem_base + FP_A_COL + 8 * a, b_desc, fp4_idesc<CG>(), sfa, sfb, t ? 1u : 0u);
        }
        umma_commit<CG>(bar);
    }
    mbar_wait_t<true>(bar, 0);
    tc_fence_after();

    uint64_t ma[4];
    for (uint32_t a = 0; a < 4; ++a) ma[a] = probe_mask(seed, 0, a, row);
    float mn = 3.0e38f, mx = -3.0e38f, top = 0.0f;
    uint32_t bad = 0;
    for (uint32_t c0 = 0; c0 < (uint32_t)FP_N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(my_lanes + FP_ACC_COL + c0, v);
        tc_wait_ld();
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            uint64_t expect = 0;
            for (uint32_t b = 0; b < 4; ++b) {
                const uint64_t mb = probe_mask(seed, 1, b, c0 + k);
                for (uint32_t a = 0; a < 4; ++a) expect += (uint64_t)times[a * 4 + b] * (uint64_t)__popcll(ma[a] & mb);
            }
            const float want = (float)expect, got = __uint_as_float(v[k]);  // expect < 2^24: exact as a float
            mn = fminf(mn, got - want); mx = fmaxf(mx, got - want); top = fmaxf(top, want);
            bad += (got != want) || expect >= (1ull << 24);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        top = fmaxf(top, __shfl_xor_sync(0xffffffffu, top, o));
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if (lane == 0) { red[warp * 4] = __float_as_uint(mn); red[warp * 4 + 1] = __float_as_uint(mx); red[warp * 4 + 2] = __float_as_uint(top); red[warp * 4 + 3] = bad; }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        Fp4RandResult r{0.0f, 3.0e38f, -3.0e38f, 0};
        for (int w = 0; w < 4; ++w) {
            r.min_diff = fminf(r.min_diff, __uint_as_float(red[w * 4]));
            r.max_diff = fmaxf(r.max_diff, __uint_as_float(red[w * 4 + 1]));
            r.max_expected = fmaxf(r.max_expected, __uint_as_float(red[w * 4 + 2]));
            r.mismatches += red[w * 4 + 3];
        }
        results[rank] = r;
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 0) tmem_free<CG>(tmem_base);
}

template <int CG>
int run_fp4_random(uint32_t n_steps, uint32_t seed, Fp4RandResult* out) {
    Fp4RandResult* d_res = nullptr;
    STORM_CUDA_TRY(cudaMalloc(&d_res, 2 * sizeof(Fp4RandResult)));
    struct Release { void* p; ~Release() { cudaFree(p); } } release{d_res};
    const int smem_bytes = 1024 + (FP_N / CG) * 128 + 512;
    STORM_CUDA_TRY(cudaFuncSetAttribute(fp4_random_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CG);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    STORM_CUDA_TRY(cudaLaunchKernelEx(&cfg, fp4_random_kernel<CG>, n_steps, seed, d_res));
    count_launch();
    Fp4RandResult h[2];
    STORM_CUDA_TRY(cudaMemcpy(h, d_res, CG * sizeof(Fp4RandResult), cudaMemcpyDeviceToHost));
    *out = h[0];
    if (CG == 2) {
        out->max_expected = fmaxf(h[0].max_expected, h[1].max_expected);
        out->min_diff = fminf(h[0].min_diff, h[1].min_diff);
        out->max_diff = fmaxf(h[0].max_diff, h[1].max_diff);
        out->mismatches = h[0].mismatches + h[1].mismatches;
    }
    return STORM_B200_OK;
}

}  // namespace

int fp4_peak_ops(int cg, double* ops_per_s, double* clock64_mhz, int n) {
    if (n == 128) return cg == 1 ? run_fp4_peak<1, 128>(ops_per_s, clock64_mhz) : run_fp4_peak<2, 128>(ops_per_s, clock64_mhz);
    return cg == 1 ? run_fp4_peak<1, FP_N>(ops_per_s, clock64_mhz) : run_fp4_peak<2, FP_N>(ops_per_s, clock64_mhz);
}

static int run_fp4_cases(const Fp4Case* cases, uint32_t n_cases, Fp4Result* results) {
    Fp4Case* d_cases = nullptr; Fp4Result* d_res = nullptr;
    STORM_CUDA_TRY(cudaMalloc(&d_cases, n_cases * sizeof(Fp4Case)));
    STORM_CUDA_TRY(cudaMalloc(&d_res, n_cases * sizeof(Fp4Result)));
    { int crc = copy_to_device_now(d_cases, cases, n_cases * sizeof(Fp4Case)); if (crc) return crc; }
    const int smem_bytes = 1024 + 2 * FP_N * 128 + 256;
    STORM_CUDA_TRY(cudaFuncSetAttribute(fp4_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    fp4_exact_kernel<<<1, 128, smem_bytes>>>(d_cases, d_res, n_cases);
    count_launch();
    STORM_CUDA_TRY(cudaGetLastError());
    STORM_CUDA_TRY(cudaMemcpy(results, d_res, n_cases * sizeof(Fp4Result), cudaMemcpyDeviceToHost));
    cudaFree(d_cases); cudaFree(d_res);
    return STORM_B200_OK;
}

// One-time check per device that the tensor core's fp32 accumulation of E2M1 products is exact for
// the integer sums the FP4 tile kernel produces: accumulators driven up to 2^24 - 1 in steps of 64
// and of 1, with every operand encoding the kernel uses.  AUTO only picks the FP4 form on a device
// that passed; 1 = exact, 0 = not (or the probe could not run).
int fp4_selftest_ok() {
    constexpr int MAX_DEV = 16;
    static int state[MAX_DEV] = {};                 // 0 unknown, 1 exact, 2 inexact
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= MAX_DEV) return 0;
    std::lock_guard<std::mutex> lock(mu);
    if (state[dev] == 0) {
        std::vector<Fp4Case> cases;
        for (uint32_t p = 0; p < 4; ++p) {
            cases.push_back({1u, 1u, p});
            cases.push_back({2048u, 3u, p});
            cases.push_back({32767u, 63u, p});       // 2^21 - 1
            cases.push_back({262143u, 63u, p});      // 2^24 - 1: the largest count the kernel accepts
        }
        std::vector<Fp4Result> res(cases.size());
        bool ok = run_fp4_cases(cases.data(), (uint32_t)cases.size(), res.data()) == STORM_B200_OK;
        for (const Fp4Result& r : res) ok = ok && r.mismatches == 0;
        // data-dependent increments (0 .. 64 per instruction and element), both cta_group forms, up to 2^24 - 64
        for (uint32_t steps : {1000u, 262143u}) {
            Fp4RandResult r1{}, r2{};
            ok = ok && run_fp4_random<1>(steps, 17u + steps, &r1) == STORM_B200_OK && r1.mismatches == 0 && r1.max_expected == 64.0f * (float)steps;
            ok = ok && run_fp4_random<2>(steps, 29u + steps, &r2) == STORM_B200_OK && r2.mismatches == 0 && r2.max_expected == 64.0f * (float)steps;
        }
        state[dev] = ok ? 1 : 2;
    }
    return state[dev] == 1;
}

}  // namespace storm

// cases: n_cases x {n_full, n_single, pattern}; results: n_cases x {expect, min, max (as float), mismatches (u32 bits)}.
extern "C" int STORM_b200_fp4_probe(const uint32_t* cases, uint32_t n_cases, float* results) {
    using namespace storm;
    int rc = require_device();
    if (rc) return rc;
    if (!cases || !results || n_cases == 0) { set_error("fp4 probe: bad arguments"); return STORM_B200_EINVAL; }
    static_assert(sizeof(Fp4Case) == 12 && sizeof(Fp4Result) == 16, "C-ABI layout of the probe records");
    return run_fp4_cases(reinterpret_cast<const Fp4Case*>(cases), n_cases, reinterpret_cast<Fp4Result*>(results));
}

// Data-dependent increments (see fp4_random_kernel): n_steps instructions (at most 262 143: the all-ones element then holds
// 2^24 - 64) at cta_group cg (1 or 2).  results: {largest expected element, smallest and largest (got - expected)} as floats
// and the number of elements that differ (u32 bits).
extern "C" int STORM_b200_fp4_probe_random(int cg, uint32_t n_steps, uint32_t seed, float* results) {
    using namespace storm;
    int rc = require_device();
    if (rc) return rc;
    if (!results || n_steps == 0 || n_steps > 262143u || (cg != 1 && cg != 2)) { set_error("fp4 random probe: bad arguments"); return STORM_B200_EINVAL; }
    static_assert(sizeof(Fp4RandResult) == 16, "C-ABI layout of the probe record");
    Fp4RandResult r{};
    rc = cg == 1 ? run_fp4_random<1>(n_steps, seed, &r) : run_fp4_random<2>(n_steps, seed, &r);
    if (rc) return rc;
    memcpy(results, &r, sizeof(r));
    return STORM_B200_OK;
}

extern "C" int STORM_b200_fp4_selftest(void) {
    using namespace storm;
    if (require_device()) return 0;
    return fp4_selftest_ok();
}
