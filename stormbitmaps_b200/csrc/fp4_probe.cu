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
template <int CG>
__host__ __device__ constexpr uint32_t fp4_idesc() {
    return (1u << 7) | (1u << 10) | ((uint32_t)(FP_N >> 3) << 17) | (1u << 23) | ((uint32_t)((128 * CG) >> 4) << 24);
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

template <int CG>
__global__ void __launch_bounds__(128, 1) fp4_peak_kernel(uint32_t iters) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    constexpr uint32_t RING = 4;
    constexpr uint32_t B_BYTES = (FP_N / CG) * 128;
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
    if (rank == 0 && tid == 0) {
        const uint64_t desc_hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t sfa = tmem_base + FP_SF_COL, sfb = tmem_base + FP_SF_COL + 32;
        for (uint32_t it = 0; it < iters; ++it) {
            if (it >= RING) mbar_wait(bar_base + 8 * (it % RING), ((it / RING) - 1) & 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t b_desc = desc_hi | (uint64_t)(((smem_base + k * 32) >> 4) & 0x3FFF);
                umma_mxf4_ts<CG>(tmem_base + FP_ACC_COL, tmem_base + FP_A_COL + k * 8, b_desc, fp4_idesc<CG>(), sfa, sfb, 1u);
            }
            umma_commit<CG>(bar_base + 8 * (it % RING));
        }
        for (uint32_t it = iters > RING ? iters - RING : 0; it < iters; ++it)
            mbar_wait(bar_base + 8 * (it % RING), (it / RING) & 1);
    }
    __syncwarp();
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 0) tmem_free<CG>(tmem_base);
}

template <int CG>
int run_fp4_peak(double* ops_per_s) {
    int dev = 0, sms = 0;
    STORM_CUDA_TRY(cudaGetDevice(&dev));
    STORM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int smem_bytes = 1024 + (FP_N / CG) * 128 + 256;
    STORM_CUDA_TRY(cudaFuncSetAttribute(fp4_peak_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
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
    const uint32_t iters = 100000;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {                                   // rep 0 is the warm-up
        STORM_CUDA_TRY(cudaEventRecord(e0));
        STORM_CUDA_TRY(cudaLaunchKernelEx(&cfg, fp4_peak_kernel<CG>, iters));
        STORM_CUDA_TRY(cudaEventRecord(e1));
        STORM_CUDA_TRY(cudaEventSynchronize(e1));
        count_launch();
        float ms = 0;
        STORM_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        // per SM and instruction: 128 x 256 x 64 MACs = 2 ops each
        const double ops = (double)cfg.gridDim.x * iters * 4.0 * 128.0 * 256.0 * 64.0 * 2.0;
        if (rep > 0 && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ops_per_s = best;
    return STORM_B200_OK;
}

}  // namespace

int fp4_peak_ops(int cg, double* ops_per_s) { return cg == 1 ? run_fp4_peak<1>(ops_per_s) : run_fp4_peak<2>(ops_per_s); }

static int run_fp4_cases(const Fp4Case* cases, uint32_t n_cases, Fp4Result* results) {
    Fp4Case* d_cases = nullptr; Fp4Result* d_res = nullptr;
    STORM_CUDA_TRY(cudaMalloc(&d_cases, n_cases * sizeof(Fp4Case)));
    STORM_CUDA_TRY(cudaMalloc(&d_res, n_cases * sizeof(Fp4Result)));
    STORM_CUDA_TRY(cudaMemcpy(d_cases, cases, n_cases * sizeof(Fp4Case), cudaMemcpyHostToDevice));
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

extern "C" int STORM_b200_fp4_selftest(void) {
    using namespace storm;
    if (require_device()) return 0;
    return fp4_selftest_ok();
}
