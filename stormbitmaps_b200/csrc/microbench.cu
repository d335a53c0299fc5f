// microbench.cu -- issue-rate micro-benchmarks for the CUDA-core roofline.
//
// MEASURED_PEAKS.json carries only the HBM and bf16 peaks; the CUDA-core tile
// kernels are bound by the POPC (and, for the carry-save variant, the ALU) pipe,
// so their roofline denominators are measured here on the device itself
// (SURVEY.md section 7.3 item 1).  Each kernel runs 8 independent dependency
// chains per thread, ONE CTA of 16 warps per SM, long enough to amortise the launch.
// (Round 1 launched two CTAs of 8 warps per SM and divided by the clock64 delta of CTA 0, which only
// spans its own half of the kernel: every per-clock figure came out 2x too high.  One clock64 tick is one
// SM cycle -- kind 8 below measures 1964.4 ticks per microsecond at nvidia-smi's 1965 MHz.)
#include <algorithm>

#include "common.cuh"

namespace storm {
namespace {

constexpr int MB_THREADS = 512;
constexpr int MB_UNROLL = 16;

template <int KIND>
__global__ void __launch_bounds__(MB_THREADS) mb_kernel(uint32_t* out, int iters, uint32_t seed, long long* cycles) {
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = seed * 2654435761u + threadIdx.x * 97u + i * 7919u + blockIdx.x;
    const uint32_t a = seed | 1u, b = ~seed;
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < MB_UNROLL; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (KIND == 0) {
                    asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
                } else if (KIND == 1) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
                } else if (KIND == 2) {
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
                } else {   // 1 POPC : 2 LOP3, the direct AND+popcount mix
                    if ((i & 3) == 0) asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
                    else if ((i & 3) != 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
                }
            }
        }
    }
    const long long c1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= x[i];
    if (r == 0x12345678u) out[0] = r;                 // keep the chains alive
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = c1 - c0;
}

template <int KIND>
int run_kind(double* rate, double* mhz, double ops_per_inner) {
    int dev = 0, sms = 0;
    STORM_CUDA_TRY(cudaGetDevice(&dev));
    STORM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint32_t* d_out = nullptr; long long* d_cyc = nullptr;
    STORM_CUDA_TRY(cudaMalloc(&d_out, 4));
    STORM_CUDA_TRY(cudaMalloc(&d_cyc, 8));
    cudaEvent_t e0, e1;
    STORM_CUDA_TRY(cudaEventCreate(&e0));
    STORM_CUDA_TRY(cudaEventCreate(&e1));
    const int grid = sms, iters = 4096;
    double best = 0, best_mhz = 0;
    for (int rep = 0; rep < 4; ++rep) {                // rep 0 is the warm-up
        STORM_CUDA_TRY(cudaEventRecord(e0));
        mb_kernel<KIND><<<grid, MB_THREADS>>>(d_out, iters, 12345u + rep, d_cyc);
        STORM_CUDA_TRY(cudaEventRecord(e1));
        STORM_CUDA_TRY(cudaEventSynchronize(e1));
        count_launch();
        float ms = 0;
        STORM_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        long long cyc = 0;
        STORM_CUDA_TRY(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
        const double ops = (double)grid * MB_THREADS * iters * MB_UNROLL * ops_per_inner;
        const double r = ops / (ms * 1e-3);
        if (rep > 0 && r > best) { best = r; best_mhz = (double)cyc / (ms * 1e-3) / 1e6; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out); cudaFree(d_cyc);
    *rate = best;
    if (mhz) *mhz = best_mhz;
    return STORM_B200_OK;
}

// How fast does clock64() tick?  Every per-clock figure of this library divides by a clock64 delta, and the first
// round's table was off by a factor of two against nvidia-smi's SM clock.  One thread per SM spins for 250 ms of
// %globaltimer and reports its clock64 delta: ticks per nanosecond, to be read beside `nvidia-smi clocks.sm`
// sampled during the spin (the part idles at its maximum clock here: nothing else runs).
__global__ void clock_calibration_kernel(unsigned long long spin_ns, unsigned long long* out) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < spin_ns);
    const long long c1 = clock64();
    out[2 * blockIdx.x] = (unsigned long long)(c1 - c0);
    out[2 * blockIdx.x + 1] = t1 - t0;
}

int run_clock_calibration(double* ticks_per_s, double* mhz) {
    int dev = 0, sms = 0;
    STORM_CUDA_TRY(cudaGetDevice(&dev));
    STORM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    unsigned long long* d = nullptr;
    STORM_CUDA_TRY(cudaMalloc(&d, 2 * sms * sizeof(unsigned long long)));
    clock_calibration_kernel<<<sms, 1>>>(250000000ull, d);
    count_launch();
    unsigned long long h[2 * 256] = {};
    const cudaError_t e = cudaMemcpy(h, d, 2 * std::min(sms, 256) * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    STORM_CUDA_TRY(e);
    double c = 0, t = 0;
    for (int i = 0; i < std::min(sms, 256); ++i) { c += (double)h[2 * i]; t += (double)h[2 * i + 1]; }
    *ticks_per_s = c / t * 1e9;
    if (mhz) *mhz = c / t * 1e3;
    return STORM_B200_OK;
}

}  // namespace
}  // namespace storm

extern "C" int STORM_b200_microbench(int kind, double* rate, double* sm_mhz) {
    using namespace storm;
    if (!rate) { set_error("rate is NULL"); return STORM_B200_EINVAL; }
    switch (kind) {
        case 0: return run_kind<0>(rate, sm_mhz, 8.0);
        case 1: return run_kind<1>(rate, sm_mhz, 8.0);
        case 2: return run_kind<2>(rate, sm_mhz, 8.0);
        case 3: return run_kind<3>(rate, sm_mhz, 6.0);   // 2 POPC + 4 LOP3 per inner step
        case 4: return umma_peak_ops(1, rate, sm_mhz);   // tcgen05.mma kind::i8, cta_group::1
        case 5: return umma_peak_ops(2, rate, sm_mhz);   // tcgen05.mma kind::i8, cta_group::2
        case 6: return fp4_peak_ops(1, rate, sm_mhz);    // tcgen05.mma kind::mxf4 (E2M1, K 64), cta_group::1
        case 7: return fp4_peak_ops(2, rate, sm_mhz);    // tcgen05.mma kind::mxf4, cta_group::2
        case 9: return fp4_peak_ops(2, rate, sm_mhz, 128);   // kind::mxf4, cta_group::2, N = 128 (the per-pair form's instruction)
        case 8: return run_clock_calibration(rate, sm_mhz);   // clock64 ticks per second (and per 1e6) over a 250 ms spin
        default: set_error("unknown microbench kind %d", kind); return STORM_B200_EINVAL;
    }
}
