// synth.cu -- device-side row construction.
//
//  * synthetic inputs for bench/tests, bit-identical to the CPU generator in
//    oracle/storm_oracle.c (recipe: benchmark.cpp:749-797 -- per row, n_draws
//    uniform positions WITH replacement, duplicates collapse);
//  * scatter_positions: the device half of STORM_contig_add / add_bulk
//    (storm.c:1103-1115: data[v/64] |= 1 << (v%64)).
#include "common.cuh"

namespace storm {

namespace {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline uint64_t row_key(uint64_t seed, uint64_t row) {
    return splitmix64(splitmix64(seed) ^ (row * 0xD1342543DE82EF95ull));
}
__host__ __device__ inline uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

// one thread per (row, draw)
__global__ void synth_uniform_kernel(uint64_t* rows, uint64_t n_rows, uint64_t stride, uint32_t M,
                                     uint32_t n_draws, uint64_t seed, uint64_t row0) {
    const uint64_t total = n_rows * (uint64_t)n_draws;
    for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = idx / n_draws, t = idx % n_draws;
        const uint64_t z = splitmix64(row_key(seed, row0 + r) + t * 0x9E3779B97F4A7C15ull);
        const uint32_t p = (uint32_t)(((z >> 32) * (uint64_t)M) >> 32);
        atomicOr(reinterpret_cast<unsigned long long*>(rows + r * stride + (p >> 6)), 1ull << (p & 63));
    }
}

// one thread per (row, word)
__global__ void synth_geno_kernel(uint64_t* rows, uint64_t n_rows, uint64_t stride, uint32_t M,
                                  uint64_t seed, uint64_t row0) {
    const uint32_t n_words = (M + 63) / 64;
    const uint64_t total = n_rows * (uint64_t)n_words;
    for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = idx / n_words;
        const uint32_t w = (uint32_t)(idx % n_words);
        const uint64_t u24 = row_key(seed ^ 0x47454E4Full, row0 + r) >> 40;
        uint64_t thr64 = (u24 * u24) >> 17;
        const uint32_t thr = (uint32_t)(thr64 < 21474836ull ? 21474836ull : thr64);
        const uint32_t key = (uint32_t)row_key(seed, row0 + r);
        uint64_t word = 0;
        const uint32_t k0 = w * 64;
#pragma unroll 8
        for (uint32_t b = 0; b < 64; ++b) {
            const uint32_t k = k0 + b;
            if (k < M && fmix32(k * 0x9E3779B1u + key) < thr) word |= 1ull << b;
        }
        rows[r * stride + w] = word;
    }
}

// one warp per row: positions[off[r] .. off[r+1]) -> bits of row r
__global__ void scatter_positions_kernel(uint64_t* rows, uint64_t stride, const uint32_t* pos,
                                         const uint64_t* off, uint64_t n_rows) {
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n_rows; r += n_warps) {
        const uint64_t b = off[r], e = off[r + 1];
        for (uint64_t k = b + lane; k < e; k += 32) {
            const uint32_t p = pos[k];
            atomicOr(reinterpret_cast<unsigned long long*>(rows + r * stride + (p >> 6)), 1ull << (p & 63));
        }
    }
}

inline unsigned grid_for(uint64_t work, unsigned block, unsigned cap = 148 * 32) {
    uint64_t g = (work + block - 1) / block;
    if (g < 1) g = 1;
    return (unsigned)(g > cap ? cap : g);
}

}  // namespace

int launch_synth_uniform(uint64_t* d_rows, uint64_t n_rows, uint64_t stride, uint32_t M,
                         uint32_t n_draws, uint64_t seed, uint64_t row0, cudaStream_t stream) {
    if (n_rows == 0 || n_draws == 0) return STORM_B200_OK;
    synth_uniform_kernel<<<grid_for(n_rows * (uint64_t)n_draws, 256), 256, 0, stream>>>(
        d_rows, n_rows, stride, M, n_draws, seed, row0);
    STORM_CUDA_TRY(cudaGetLastError());
    count_launch();
    return STORM_B200_OK;
}

int launch_synth_geno(uint64_t* d_rows, uint64_t n_rows, uint64_t stride, uint32_t M,
                      uint64_t seed, uint64_t row0, cudaStream_t stream) {
    if (n_rows == 0) return STORM_B200_OK;
    synth_geno_kernel<<<grid_for(n_rows * (uint64_t)((M + 63) / 64), 256), 256, 0, stream>>>(
        d_rows, n_rows, stride, M, seed, row0);
    STORM_CUDA_TRY(cudaGetLastError());
    count_launch();
    return STORM_B200_OK;
}

int launch_scatter_positions(uint64_t* d_rows, uint64_t stride, const uint32_t* d_pos,
                             const uint64_t* d_off, uint64_t n_rows, cudaStream_t stream) {
    if (n_rows == 0) return STORM_B200_OK;
    scatter_positions_kernel<<<grid_for(n_rows * 32, 256), 256, 0, stream>>>(d_rows, stride, d_pos, d_off, n_rows);
    STORM_CUDA_TRY(cudaGetLastError());
    count_launch();
    return STORM_B200_OK;
}

}  // namespace storm
