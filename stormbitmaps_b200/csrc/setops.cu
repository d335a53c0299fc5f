// setops.cu -- union and symmetric-difference cardinalities on top of the intersection tiles.
//
// libalgebra ships three per-pair kernel families with one signature (libalgebra.h:3035):
// intersect = popcount(a & b), union = popcount(a | b) (libalgebra.h:2994-3000, 521-540) and
// diff = popcount(a ^ b) (libalgebra.h:3002-3008, 543-563), each with its own CPUID chooser
// (libalgebra.h:3094-3140, 3142-3188, 3190-3236), and storm.c's raw-buffer loops take any of them
// as `f` (storm.c:132-150).  On the GPU the two extra families need no kernel of their own:
//
//     |a | b| = |a| + |b| -     |a & b|            |a ^ b| = |a| + |b| - 2 |a & b|
//
// so one pass of row popcounts (HBM-bound, 8 N W bytes) plus the intersection tiles answers them,
// for totals and for per-pair rectangles alike.
#include <algorithm>

#include "common.cuh"
#include "runtime.h"

namespace storm {
namespace {

// one warp per row: 16-byte loads, POPC, shuffle reduction
__global__ void __launch_bounds__(256) row_popcount_kernel(const uint64_t* __restrict__ rows, uint64_t n_rows, uint32_t n_words,
                                                           uint64_t stride, uint32_t* __restrict__ counts) {
    const uint64_t row = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t* p = rows + row * stride;
    uint32_t c = 0;
    const uint32_t n2 = n_words >> 1;                                      // rows are 16-byte aligned (even stride)
    const uint4* p4 = reinterpret_cast<const uint4*>(p);
    for (uint32_t k = lane; k < n2; k += 32) {
        const uint4 v = p4[k];
        c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
    if ((n_words & 1) && lane == 0) c += __popcll(p[n_words - 1]);
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) counts[row] = c;
}

// Sum over the unmasked pairs (i, j) of the rectangle of |a_i| + |b_j|: every A row counts once per
// valid partner j, every B row once per valid partner i.  Global indices gi = i_off + i, gj = j_off + j;
// with strict_upper only pairs with gj > gi exist.
__global__ void __launch_bounds__(256) pop_term_kernel(const uint32_t* __restrict__ ca, uint64_t nA, uint64_t i_off,
                                                       const uint32_t* __restrict__ cb, uint64_t nB, uint64_t j_off,
                                                       int strict_upper, unsigned long long* __restrict__ term) {
    unsigned long long acc = 0;
    const uint64_t n = nA + nB;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        if (t < nA) {
            uint64_t partners = nB;
            if (strict_upper) {                                            // gj in (gi, j_off + nB)
                const uint64_t gi = i_off + t, lo = gi + 1 > j_off ? gi + 1 : j_off, hi = j_off + nB;
                partners = hi > lo ? hi - lo : 0;
            }
            acc += (unsigned long long)ca[t] * partners;
        } else {
            const uint64_t j = t - nA;
            uint64_t partners = nA;
            if (strict_upper) {                                            // gi in [i_off, min(i_off + nA, gj))
                const uint64_t gj = j_off + j, hi = gj < i_off + nA ? gj : i_off + nA;
                partners = hi > i_off ? hi - i_off : 0;
            }
            acc += (unsigned long long)cb[j] * partners;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(term, acc);
}

// *total += term - k * intersect
__global__ void op_finish_kernel(const unsigned long long* term, const unsigned long long* isect, unsigned long long k,
                                 unsigned long long* total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(total, *term - k * *isect);
}

// out[i][j] (intersection counts, masked entries already 0) -> |a_i| + |b_j| - k out[i][j] on unmasked entries
__global__ void __launch_bounds__(256) op_pairs_kernel(uint32_t* __restrict__ out, uint64_t ld, uint64_t nA, uint64_t nB,
                                                       const uint32_t* __restrict__ ca, const uint32_t* __restrict__ cb,
                                                       uint64_t i_off, uint64_t j_off, int strict_upper, uint32_t k) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t i = blockIdx.y;
    if (j >= nB || i >= nA) return;
    if (strict_upper && j_off + j <= i_off + i) return;
    uint32_t* p = out + i * ld + j;
    *p = ca[i] + cb[j] - k * *p;
}

}  // namespace

int launch_row_popcounts(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words, uint64_t stride, uint32_t* d_counts,
                         cudaStream_t stream) {
    if (n_rows == 0) return STORM_B200_OK;
    row_popcount_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, stream>>>(d_rows, n_rows, n_words, stride, d_counts);
    count_launch();
    STORM_CUDA_TRY(cudaGetLastError());
    return STORM_B200_OK;
}

// Rectangle (or, with A == B and offsets 0 and strict_upper, the whole triangle) under a set operation.
// Intersections come from the regular tile kernels (`shard` of `n_shards` in triangle mode is not offered
// here: the popcount term is not shardable by tiles; multi-GPU callers combine it themselves).
int pairw_rect_op(const uint64_t* dA, uint64_t nA, uint64_t strideA, uint64_t i_off,
                  const uint64_t* dB, uint64_t nB, uint64_t strideB, uint64_t j_off,
                  uint32_t n_words, int strict_upper, int op, int kernel, bool triangle,
                  uint32_t* d_out, uint64_t ld, uint64_t* d_total, cudaStream_t stream) {
    if (op == STORM_B200_OP_INTERSECT) {
        if (triangle) return pairw_triangle(dA, nA, n_words, strideA, 0, 1, kernel, d_total, stream);
        return pairw_rect(dA, nA, strideA, i_off, dB, nB, strideB, j_off, n_words, strict_upper, kernel, d_out, ld, d_total, stream);
    }
    if (op != STORM_B200_OP_UNION && op != STORM_B200_OP_DIFF) { set_error("unknown set operation %d", op); return STORM_B200_EINVAL; }
    int rc = require_device();
    if (rc) return rc;
    if (nA == 0 || nB == 0) return STORM_B200_OK;
    if ((rc = check_rows(dA, strideA, n_words)) || (rc = check_rows(dB, strideB, n_words))) return rc;   // before any launch: the popcount kernel loads 16 bytes at a time
    const unsigned long long k = op == STORM_B200_OP_UNION ? 1 : 2;
    const bool same = dA == dB && nA == nB && strideA == strideB;
    // one stream-ordered allocation: [2 x u64: popcount term, intersections][row popcounts of A (and B)]
    unsigned long long* scal = nullptr;
    STORM_CUDA_TRY(cudaMallocAsync(&scal, 2 * sizeof(unsigned long long) + (nA + (same ? 0 : nB)) * sizeof(uint32_t), stream));
    struct Release { void* p; cudaStream_t s; ~Release() { cudaFreeAsync(p, s); } } release{scal, stream};   // every return path
    uint32_t* counts = reinterpret_cast<uint32_t*>(scal + 2);
    STORM_CUDA_TRY(cudaMemsetAsync(scal, 0, 2 * sizeof(unsigned long long), stream));
    uint32_t* ca = counts;
    uint32_t* cb = same ? counts : counts + nA;
    rc = launch_row_popcounts(dA, nA, n_words, strideA, ca, stream);
    if (!rc && !same) rc = launch_row_popcounts(dB, nB, n_words, strideB, cb, stream);
    if (!rc) {
        if (triangle) rc = pairw_triangle(dA, nA, n_words, strideA, 0, 1, kernel, reinterpret_cast<uint64_t*>(scal + 1), stream);
        else rc = pairw_rect(dA, nA, strideA, i_off, dB, nB, strideB, j_off, n_words, strict_upper, kernel, d_out, ld,
                             d_total ? reinterpret_cast<uint64_t*>(scal + 1) : nullptr, stream);
    }
    if (!rc && d_total) {
        const uint64_t n = nA + nB;
        pop_term_kernel<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 1184), 256, 0, stream>>>(ca, nA, i_off, cb, nB, j_off,
                                                                                                strict_upper, scal);
        op_finish_kernel<<<1, 32, 0, stream>>>(scal, scal + 1, k, reinterpret_cast<unsigned long long*>(d_total));
        count_launch(2);
        if (cudaGetLastError() != cudaSuccess) rc = STORM_B200_ECUDA;
    }
    if (!rc && d_out) {
        dim3 grid((unsigned)((nB + 255) / 256), (unsigned)nA);
        if (nA > 65535) { set_error("per-pair set-operation rectangles are limited to 65535 rows per call"); rc = STORM_B200_EINVAL; }
        else {
            op_pairs_kernel<<<grid, 256, 0, stream>>>(d_out, ld, nA, nB, ca, cb, i_off, j_off, strict_upper, (uint32_t)k);
            count_launch();
            if (cudaGetLastError() != cudaSuccess) rc = STORM_B200_ECUDA;
        }
    }
    return rc;
}

}  // namespace storm

using namespace storm;

extern "C" {

int STORM_b200_row_popcounts_device(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words, uint64_t row_stride_words,
                                    uint32_t* d_counts, void* stream) {
    int rc = require_device();
    if (rc) return rc;
    if (!d_rows || !d_counts || n_words == 0 || n_words > row_stride_words || (row_stride_words & 1) || ((uintptr_t)d_rows & 15)) {
        set_error("row popcounts: bad arguments");
        return STORM_B200_EINVAL;
    }
    return launch_row_popcounts(d_rows, n_rows, n_words, row_stride_words, d_counts, (cudaStream_t)stream);
}

int STORM_b200_pairw_op_device(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words, uint64_t row_stride_words,
                               int op, int kernel, uint64_t* d_total, void* stream) {
    if (!d_total) { set_error("d_total is NULL"); return STORM_B200_EINVAL; }
    if (n_rows < 2) return STORM_B200_OK;
    return pairw_rect_op(d_rows, n_rows, row_stride_words, 0, d_rows, n_rows, row_stride_words, 0, n_words, 1, op, kernel, true,
                         nullptr, 0, d_total, (cudaStream_t)stream);
}

int STORM_b200_pairw_rect_op_device(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words, uint64_t row_stride_words,
                                    uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, int strict_upper, int op, int kernel,
                                    uint32_t* d_out, uint64_t ld, uint64_t* d_total, void* stream) {
    if (i0 > i1 || j0 > j1 || i1 > n_rows || j1 > n_rows) { set_error("rectangle outside %llu rows", (unsigned long long)n_rows); return STORM_B200_EINVAL; }
    return pairw_rect_op(d_rows + i0 * row_stride_words, i1 - i0, row_stride_words, i0, d_rows + j0 * row_stride_words, j1 - j0,
                         row_stride_words, j0, n_words, strict_upper, op, kernel, false, d_out, ld, d_total, (cudaStream_t)stream);
}

}  // extern "C"
