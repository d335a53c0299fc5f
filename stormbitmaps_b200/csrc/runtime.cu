// runtime.cu -- error state, launch accounting, device queries, tile bookkeeping
// and the device-buffer entry points of storm_b200.h.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "runtime.h"

namespace storm {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_default_kernel{STORM_B200_KERNEL_AUTO};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_error; }
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int default_kernel() { return g_default_kernel.load(); }

int require_device() {
    static std::atomic<int> known[64];                 // per device ordinal: 1 = an sm_100 device (checked once)
    int dev = -1;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && known[dev].load(std::memory_order_relaxed) == 1) return STORM_B200_OK;
    cudaGetLastError();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return STORM_B200_ENODEV;
    }
    STORM_CUDA_TRY(cudaGetDevice(&dev));
    int major = 0;
    STORM_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        set_error("device %d has compute capability %d.x; libstorm_b200 is built for sm_100a only", dev, major);
        return STORM_B200_ENODEV;
    }
    if (dev < 64) known[dev].store(1, std::memory_order_relaxed);
    return STORM_B200_OK;
}

// ---- tile rasters ---------------------------------------------------------------
TileShape tile_shape_for(int kernel) {
    switch (kernel) {
        case STORM_B200_KERNEL_UMMA:
        case STORM_B200_KERNEL_FP4:  return umma_tile_shape();
        case STORM_B200_KERNEL_CSA:  return csa_tile_shape();
        case STORM_B200_KERNEL_B1:   return b1_tile_shape();
        default:                     return popc_tile_shape();
    }
}

uint64_t triangle_prefix(uint64_t n_rows, TileShape ts, std::vector<uint64_t>* prefix, uint32_t* n_bi, uint32_t* n_bj) {
    const uint32_t nbi = (uint32_t)((n_rows + ts.tm - 1) / ts.tm);
    const uint32_t nbj = (uint32_t)((n_rows + ts.tn - 1) / ts.tn);
    const uint32_t n_groups = (nbj + TRI_GROUP - 1) / TRI_GROUP;
    if (prefix) prefix->assign((size_t)n_groups + 1, 0);
    uint64_t acc = 0;
    for (uint32_t g = 0; g < n_groups; ++g) {
        if (prefix) (*prefix)[g] = acc;
        acc += tri_group_tiles(g * TRI_GROUP, nbi, nbj, ts.tm, ts.tn);
    }
    if (prefix) (*prefix)[n_groups] = acc;
    if (n_bi) *n_bi = nbi;
    if (n_bj) *n_bj = nbj;
    return acc;
}

// A tiny per-thread cache of device prefix arrays keyed by (n_rows, tm, tn, device).
struct PrefixEntry { uint64_t n_rows; uint32_t tm, tn; int dev; uint64_t* d_prefix; uint64_t n_tiles; uint32_t nbi, nbj; };
// Host -> device copy that is COMPLETE on return.  cudaMemcpy from pageable memory returns once the source has been
// staged, not once the DMA has reached the device; the kernels that read these tables run on non-blocking streams
// (no implicit ordering with the legacy stream the copy used) and possibly from another host thread, so wait for it.
int copy_to_device_now(void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return STORM_B200_OK;
    STORM_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    STORM_CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));
    return STORM_B200_OK;
}

static std::mutex g_prefix_mu;
static std::vector<PrefixEntry> g_prefix_cache;

int get_triangle_prefix(uint64_t n_rows, TileShape ts, cudaStream_t stream, const uint64_t** d_prefix,
                        uint64_t* n_tiles, uint32_t* nbi, uint32_t* nbj) {
    int dev = 0;
    STORM_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_prefix_mu);
    for (const PrefixEntry& e : g_prefix_cache)
        if (e.n_rows == n_rows && e.tm == ts.tm && e.tn == ts.tn && e.dev == dev) {
            *d_prefix = e.d_prefix; *n_tiles = e.n_tiles; *nbi = e.nbi; *nbj = e.nbj;
            return STORM_B200_OK;
        }
    std::vector<uint64_t> h;
    PrefixEntry e{n_rows, ts.tm, ts.tn, dev, nullptr, 0, 0, 0};
    e.n_tiles = triangle_prefix(n_rows, ts, &h, &e.nbi, &e.nbj);
    STORM_CUDA_TRY(cudaMalloc(&e.d_prefix, h.size() * sizeof(uint64_t)));
    // complete before the entry is published: other threads launch on other streams right after a cache hit
    { int rc = copy_to_device_now(e.d_prefix, h.data(), h.size() * sizeof(uint64_t)); if (rc) { cudaFree(e.d_prefix); return rc; } }
    (void)stream;
    if (g_prefix_cache.size() >= 64) {              // bounded: drop the oldest
        cudaFree(g_prefix_cache.front().d_prefix);
        g_prefix_cache.erase(g_prefix_cache.begin());
    }
    g_prefix_cache.push_back(e);
    *d_prefix = e.d_prefix; *n_tiles = e.n_tiles; *nbi = e.nbi; *nbj = e.nbj;
    return STORM_B200_OK;
}

void shard_range(uint64_t n_tiles, uint32_t shard, uint32_t n_shards, uint64_t* begin, uint64_t* end) {
    // contiguous, sizes differ by at most one tile
    const uint64_t q = n_tiles / n_shards, r = n_tiles % n_shards;
    *begin = shard * q + (shard < r ? shard : r);
    *end = *begin + q + (shard < r ? 1 : 0);
}

int resolve_kernel(int kernel, const DenseJob& job) {
    if (kernel == STORM_B200_KERNEL_AUTO) kernel = default_kernel();
    if (kernel == STORM_B200_KERNEL_AUTO) {
        // tensor-core forms first: FP4 (twice the rate of i8) where every pair count stays fp32-exact and the
        // device passed the accumulation self-test, else i8; CUDA cores for shapes the TMA path cannot take
        if (umma_fp4_supports(job) && fp4_selftest_ok()) kernel = STORM_B200_KERNEL_FP4;
        else kernel = umma_supports(job) ? STORM_B200_KERNEL_UMMA : STORM_B200_KERNEL_POPC;
    }
    return kernel;
}

int launch_dense(int kernel, const DenseJob& job, cudaStream_t stream) {
    switch (kernel) {
        case STORM_B200_KERNEL_POPC: return launch_dense_popc(job, stream);
        case STORM_B200_KERNEL_CSA:  return launch_dense_csa(job, stream);
        case STORM_B200_KERNEL_B1:   return launch_dense_b1(job, stream);
        case STORM_B200_KERNEL_UMMA:
            if (!umma_supports(job)) {
                set_error("UMMA kernel needs n_words >= 2, 16-byte aligned rows and an even row stride");
                return STORM_B200_EINVAL;
            }
            return launch_dense_umma(job, stream);
        case STORM_B200_KERNEL_FP4:
            if (!umma_fp4_supports(job)) {
                set_error("FP4 kernel needs what the UMMA kernel needs and n_words <= 262144 (pair counts below 2^24)");
                return STORM_B200_EINVAL;
            }
            return launch_dense_fp4(job, stream);
        default:
            set_error("unknown kernel id %d", kernel);
            return STORM_B200_EINVAL;
    }
}

int check_rows(const uint64_t* d_rows, uint64_t stride, uint32_t n_words) {
    if (!d_rows) { set_error("d_rows is NULL"); return STORM_B200_EINVAL; }
    if (((uintptr_t)d_rows & 15) || (stride & 1)) {
        set_error("rows must be 16-byte aligned with an even word stride (got %p, stride %llu)", (const void*)d_rows,
                  (unsigned long long)stride);
        return STORM_B200_EINVAL;
    }
    if (n_words == 0 || n_words > stride) { set_error("n_words %u out of range for stride %llu", n_words, (unsigned long long)stride); return STORM_B200_EINVAL; }
    if (n_words >= (1u << 26)) { set_error("n_words %u: per-pair counts would overflow 32 bits", n_words); return STORM_B200_EINVAL; }
    return STORM_B200_OK;
}

static DenseJob triangle_job(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words, uint64_t stride, uint64_t* d_total) {
    DenseJob job{};
    job.A = job.B = d_rows;
    job.strideA = job.strideB = stride;
    job.nA = job.nB = n_rows;
    job.n_words = n_words;
    job.strict_upper = 1;
    job.triangle = 1;
    job.total = reinterpret_cast<unsigned long long*>(d_total);
    job.reserved_sms = -1;
    return job;
}

int resolve_kernel_for_rows(int kernel, const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words, uint64_t stride) {
    return resolve_kernel(kernel, triangle_job(d_rows, n_rows, n_words, stride, nullptr));
}

// shard < n_shards: that shard's range of the raster; n_shards == 0: the explicit range [tile_begin, tile_end)
static int pairw_triangle_impl(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words, uint64_t stride,
                               uint32_t shard, uint32_t n_shards, uint64_t tile_begin, uint64_t tile_end,
                               int kernel, uint64_t* d_total, cudaStream_t stream, int reserved_sms = -1) {
    int rc = require_device();
    if (rc) return rc;
    if (!d_total) { set_error("d_total is NULL"); return STORM_B200_EINVAL; }
    if (n_rows < 2) return STORM_B200_OK;
    if ((rc = check_rows(d_rows, stride, n_words))) return rc;
    DenseJob job = triangle_job(d_rows, n_rows, n_words, stride, d_total);
    job.reserved_sms = reserved_sms;
    kernel = resolve_kernel(kernel, job);
    const TileShape ts = tile_shape_for(kernel);
    uint64_t n_tiles = 0;
    if ((rc = get_triangle_prefix(n_rows, ts, stream, &job.group_prefix, &n_tiles, &job.n_bi, &job.n_bj))) return rc;
    if (n_shards) shard_range(n_tiles, shard, n_shards, &job.tile_begin, &job.tile_end);
    else {
        if (tile_begin > tile_end || tile_end > n_tiles) { set_error("tile range [%llu, %llu) of %llu", (unsigned long long)tile_begin, (unsigned long long)tile_end, (unsigned long long)n_tiles); return STORM_B200_EINVAL; }
        job.tile_begin = tile_begin; job.tile_end = tile_end;
    }
    return launch_dense(kernel, job, stream);
}

int pairw_triangle(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words, uint64_t stride,
                   uint32_t shard, uint32_t n_shards, int kernel, uint64_t* d_total, cudaStream_t stream) {
    if (n_shards == 0 || shard >= n_shards) { set_error("shard %u of %u", shard, n_shards); return STORM_B200_EINVAL; }
    return pairw_triangle_impl(d_rows, n_rows, n_words, stride, shard, n_shards, 0, 0, kernel, d_total, stream);
}

int pairw_triangle_range(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words, uint64_t stride,
                         uint64_t tile_begin, uint64_t tile_end, int kernel, uint64_t* d_total, cudaStream_t stream, int reserved_sms) {
    return pairw_triangle_impl(d_rows, n_rows, n_words, stride, 0, 0, tile_begin, tile_end, kernel, d_total, stream, reserved_sms);
}

int pairw_rect(const uint64_t* dA, uint64_t nA, uint64_t strideA, uint64_t i_off,
               const uint64_t* dB, uint64_t nB, uint64_t strideB, uint64_t j_off,
               uint32_t n_words, int strict_upper, int kernel,
               uint32_t* d_out, uint64_t ld, uint64_t* d_total, cudaStream_t stream) {
    int rc = require_device();
    if (rc) return rc;
    if (nA == 0 || nB == 0) return STORM_B200_OK;
    if ((rc = check_rows(dA, strideA, n_words)) || (rc = check_rows(dB, strideB, n_words))) return rc;
    if (d_out && ld < nB) { set_error("ld %llu < columns %llu", (unsigned long long)ld, (unsigned long long)nB); return STORM_B200_EINVAL; }
    DenseJob job{};
    job.A = dA; job.B = dB;
    job.strideA = strideA; job.strideB = strideB;
    job.nA = nA; job.nB = nB;
    job.n_words = n_words;
    job.i_off = i_off; job.j_off = j_off;
    job.strict_upper = strict_upper;
    job.triangle = 0;
    job.out = d_out; job.ld = ld;
    job.total = reinterpret_cast<unsigned long long*>(d_total);
    job.reserved_sms = -1;
    kernel = resolve_kernel(kernel, job);
    TileShape ts = tile_shape_for(kernel);
    if (d_out && (kernel == STORM_B200_KERNEL_UMMA || kernel == STORM_B200_KERNEL_FP4)) ts = umma_pairs_tile_shape();
    job.n_bi = (uint32_t)((nA + ts.tm - 1) / ts.tm);
    job.n_bj = (uint32_t)((nB + ts.tn - 1) / ts.tn);
    job.tile_begin = 0;
    job.tile_end = (uint64_t)job.n_bi * job.n_bj;
    return launch_dense(kernel, job, stream);
}

}  // namespace storm

// =================================================================================
// C ABI
// =================================================================================
using namespace storm;

extern "C" {

const char* STORM_b200_last_error(void) { return get_error(); }
const char* STORM_b200_version(void) { return "storm-b200 0.1 (sm_100a)"; }

int STORM_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int STORM_b200_device_info(int dev, char* name, size_t name_len, int* sm_count, int* cc) {
    cudaDeviceProp p;
    STORM_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
    if (name && name_len) { strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc) *cc = p.major * 10 + p.minor;
    return STORM_B200_OK;
}

int STORM_b200_set_default_kernel(int kernel) { return g_default_kernel.exchange(kernel); }
uint64_t STORM_b200_launch_count(void) { return g_launches.load(); }

int STORM_b200_pairw_device(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                            uint64_t row_stride_words, uint32_t shard, uint32_t n_shards,
                            int kernel, uint64_t* d_total, void* stream) {
    return pairw_triangle(d_rows, n_rows, n_words, row_stride_words, shard, n_shards, kernel, d_total,
                          (cudaStream_t)stream);
}

int STORM_b200_pairw_rect_device(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                                 uint64_t row_stride_words, uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1,
                                 int strict_upper, int kernel, uint32_t* d_out, uint64_t ld, uint64_t* d_total,
                                 void* stream) {
    if (i0 > i1 || j0 > j1 || i1 > n_rows || j1 > n_rows) {
        set_error("rectangle [%llu,%llu) x [%llu,%llu) outside %llu rows", (unsigned long long)i0, (unsigned long long)i1,
                  (unsigned long long)j0, (unsigned long long)j1, (unsigned long long)n_rows);
        return STORM_B200_EINVAL;
    }
    return pairw_rect(d_rows + i0 * row_stride_words, i1 - i0, row_stride_words, i0,
                      d_rows + j0 * row_stride_words, j1 - j0, row_stride_words, j0,
                      n_words, strict_upper, kernel, d_out, ld, d_total, (cudaStream_t)stream);
}

int STORM_b200_square_device(const uint64_t* d_rows1, uint64_t n1, uint64_t stride1,
                             const uint64_t* d_rows2, uint64_t n2, uint64_t stride2,
                             uint32_t n_words, int kernel, uint32_t* d_out, uint64_t ld, uint64_t* d_total,
                             void* stream) {
    return pairw_rect(d_rows1, n1, stride1, 0, d_rows2, n2, stride2, 0, n_words, 0, kernel, d_out, ld, d_total,
                      (cudaStream_t)stream);
}

int STORM_b200_resolve_kernel(int kernel, uint32_t n_words) {
    DenseJob job{};
    job.reserved_sms = -1;
    job.n_words = n_words;
    job.strideA = job.strideB = (n_words + 15) / 16 * 16;
    return resolve_kernel(kernel, job);
}

uint64_t STORM_b200_tile_count(uint64_t n_rows, int kernel, uint32_t* tile_rows, uint32_t* tile_cols) {
    if (kernel == STORM_B200_KERNEL_AUTO) kernel = STORM_b200_resolve_kernel(kernel, 1024);
    const TileShape ts = tile_shape_for(kernel);
    if (tile_rows) *tile_rows = ts.tm;
    if (tile_cols) *tile_cols = ts.tn;
    return triangle_prefix(n_rows, ts, nullptr, nullptr, nullptr);
}

int STORM_b200_shard_tiles(uint64_t n_rows, int kernel, uint32_t shard, uint32_t n_shards,
                           uint64_t* tile_begin, uint64_t* tile_end) {
    if (n_shards == 0 || shard >= n_shards || !tile_begin || !tile_end) { set_error("shard %u of %u", shard, n_shards); return STORM_B200_EINVAL; }
    const uint64_t n_tiles = STORM_b200_tile_count(n_rows, kernel, nullptr, nullptr);
    shard_range(n_tiles, shard, n_shards, tile_begin, tile_end);
    return STORM_B200_OK;
}

int STORM_b200_tiles_below_row(uint64_t n_rows, int kernel, uint64_t row_limit, uint64_t* tile_end, uint64_t* band_rows) {
    if (!tile_end) { set_error("tile_end is NULL"); return STORM_B200_EINVAL; }
    if (kernel == STORM_B200_KERNEL_AUTO) kernel = STORM_b200_resolve_kernel(kernel, 1024);
    const TileShape ts = tile_shape_for(kernel);
    const uint64_t group_rows = (uint64_t)TRI_GROUP * ts.tn;                // rows a raster group adds
    if (band_rows) *band_rows = group_rows;
    std::vector<uint64_t> prefix;
    const uint64_t n_tiles = triangle_prefix(n_rows, ts, &prefix, nullptr, nullptr);
    // tiles of raster groups <= g only read rows below (g + 1) * group_rows (common.cuh)
    const uint64_t g = row_limit / group_rows;
    *tile_end = row_limit >= n_rows ? n_tiles : prefix[std::min<uint64_t>(g, prefix.size() - 1)];
    return STORM_B200_OK;
}

int STORM_b200_pairw_tiles_device(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                                  uint64_t row_stride_words, uint64_t tile_begin, uint64_t tile_end,
                                  int kernel, uint64_t* d_total, void* stream) {
    if (tile_begin == tile_end) return STORM_B200_OK;
    return pairw_triangle_range(d_rows, n_rows, n_words, row_stride_words, tile_begin, tile_end, kernel, d_total,
                                (cudaStream_t)stream);
}

int STORM_b200_pairw_tiles_device_ex(const uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                                     uint64_t row_stride_words, uint64_t tile_begin, uint64_t tile_end,
                                     int kernel, int reserved_sms, uint64_t* d_total, void* stream) {
    if (tile_begin == tile_end) return STORM_B200_OK;
    return pairw_triangle_range(d_rows, n_rows, n_words, row_stride_words, tile_begin, tile_end, kernel, d_total,
                                (cudaStream_t)stream, reserved_sms < 0 ? -1 : reserved_sms);
}

int STORM_b200_tile_rect(uint64_t n_rows, int kernel, uint64_t tile, uint64_t* i0, uint64_t* i1, uint64_t* j0, uint64_t* j1) {
    if (kernel == STORM_B200_KERNEL_AUTO) kernel = STORM_b200_resolve_kernel(kernel, 1024);
    const TileShape ts = tile_shape_for(kernel);
    std::vector<uint64_t> prefix;
    uint32_t nbi = 0, nbj = 0;
    const uint64_t n_tiles = triangle_prefix(n_rows, ts, &prefix, &nbi, &nbj);
    if (tile >= n_tiles || !i0 || !i1 || !j0 || !j1) { set_error("tile %llu of %llu", (unsigned long long)tile, (unsigned long long)n_tiles); return STORM_B200_EINVAL; }
    uint32_t bi = 0, bj = 0;
    tile_coords_tri(prefix.data(), nbi, nbj, tile, ts.tm, ts.tn, bi, bj);      // the function the kernels call
    *i0 = (uint64_t)bi * ts.tm; *i1 = std::min<uint64_t>(*i0 + ts.tm, n_rows);
    *j0 = (uint64_t)bj * ts.tn; *j1 = std::min<uint64_t>(*j0 + ts.tn, n_rows);
    return STORM_B200_OK;
}

int STORM_b200_synth_uniform_device(uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                                    uint64_t row_stride_words, uint32_t M, uint32_t n_draws,
                                    uint64_t seed, uint64_t row0, void* stream) {
    if (!d_rows || (uint64_t)n_words * 64 < M || n_words > row_stride_words) { set_error("bad synth arguments"); return STORM_B200_EINVAL; }
    int rc = require_device();
    if (rc) return rc;
    return launch_synth_uniform(d_rows, n_rows, row_stride_words, M, n_draws, seed, row0, (cudaStream_t)stream);
}

int STORM_b200_synth_geno_device(uint64_t* d_rows, uint64_t n_rows, uint32_t n_words,
                                 uint64_t row_stride_words, uint32_t M, uint64_t seed, uint64_t row0, void* stream) {
    if (!d_rows || (uint64_t)n_words * 64 < M || n_words > row_stride_words) { set_error("bad synth arguments"); return STORM_B200_EINVAL; }
    int rc = require_device();
    if (rc) return rc;
    return launch_synth_geno(d_rows, n_rows, row_stride_words, M, seed, row0, (cudaStream_t)stream);
}

}  // extern "C"
