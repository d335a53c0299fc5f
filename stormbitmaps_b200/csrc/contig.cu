// contig.cu -- STORM_contiguous_t: host container (reference-compatible public
// fields), device-resident 128-byte-aligned row arena, and the pairwise queries.
//
// Reference being replaced: storm.c:1001-1347 (container + four query loops) and
// storm.c:132-279 (raw-buffer wrappers).  Host code only builds and uploads the
// rows; every query is answered by CUDA kernels.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"
#include "runtime.h"

namespace storm {

namespace {

constexpr uint64_t ROW_ALIGN_WORDS = 16;   // device row stride is a multiple of 128 bytes

struct ContigState {
    int device = -1;
    cudaStream_t stream = nullptr;
    // device arena
    uint64_t* d_rows = nullptr;
    uint64_t d_cap_rows = 0;
    uint64_t stride = 0;               // words
    uint64_t uploaded_rows = 0;        // rows [0, uploaded_rows) of the host mirror are on the device
    // sparse-row position lists (host offsets + device mirrors, rebuilt lazily)
    std::vector<uint64_t> pos_off;     // per row: offset into c->scalar (valid for sparse rows)
    uint32_t* d_pos = nullptr; uint64_t d_pos_cap = 0;
    uint64_t* d_pos_off = nullptr; uint32_t* d_is_sparse = nullptr; uint32_t* d_sparse_rows = nullptr;
    uint32_t* d_dense_rows = nullptr; uint64_t d_meta_cap = 0;
    uint64_t list_rows_synced = 0;     // lists of rows [0, list_rows_synced) are on the device
    uint64_t n_sparse = 0, n_dense = 0;
    std::vector<uint32_t> h_group_start; uint32_t* d_group_start = nullptr;   // row groups of the stream kernel (all rows sparse)
    bool stream_ok = false;            // every row is sparse and short enough for the stream kernel
    int last_list_route = 0;           // 1 tile kernel, 2 tile + probe kernels, 3 stream kernel
    uint64_t* d_gather = nullptr; uint64_t d_gather_cap = 0;   // compact arena of dense rows (list path)
    // result + timing
    unsigned long long* d_total = nullptr;
    unsigned long long* h_total = nullptr;   // pinned
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    double timing[3] = {0, 0, 0};
};

struct DeviceGuard {
    int prev = -1; bool active = false;
    explicit DeviceGuard(int dev) {
        if (dev < 0) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) { cudaSetDevice(dev); active = true; }
    }
    ~DeviceGuard() { if (active) cudaSetDevice(prev); }
};

inline ContigState* state_of(STORM_contiguous_t* c) { return static_cast<ContigState*>(c->b200); }

// The per-pair host function kept in `intsec_func` so that callers that invoke the
// public field keep working ("function-pointer identity").
// It is not used by any query of this library.
uint64_t host_intersect_count(const uint64_t* a, const uint64_t* b, const size_t n) {
    uint64_t c = 0;
    for (size_t k = 0; k < n; ++k) c += (uint64_t)__builtin_popcountll(a[k] & b[k]);
    return c;
}

int ensure_device_state(ContigState* st) {
    if (st->device >= 0 && st->d_total) return STORM_B200_OK;
    int rc = require_device();
    if (rc) return rc;
    STORM_CUDA_TRY(cudaGetDevice(&st->device));
    STORM_CUDA_TRY(cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking));
    STORM_CUDA_TRY(cudaMalloc(&st->d_total, sizeof(unsigned long long)));
    STORM_CUDA_TRY(cudaMallocHost(&st->h_total, sizeof(unsigned long long)));
    for (auto& e : st->ev) STORM_CUDA_TRY(cudaEventCreate(&e));
    return STORM_B200_OK;
}

int ensure_device_rows(STORM_contiguous_t* c, ContigState* st, uint64_t rows) {
    if (st->stride == 0)
        st->stride = ((uint64_t)c->n_bitmaps_vector + ROW_ALIGN_WORDS - 1) / ROW_ALIGN_WORDS * ROW_ALIGN_WORDS;
    if (rows <= st->d_cap_rows) return STORM_B200_OK;
    uint64_t cap = std::max<uint64_t>(rows, st->d_cap_rows + st->d_cap_rows / 2);
    cap = (cap + 511) / 512 * 512;
    uint64_t* fresh = nullptr;
    if (cudaMalloc(&fresh, cap * st->stride * sizeof(uint64_t)) != cudaSuccess) {
        cudaGetLastError();
        set_error("device arena of %llu rows x %llu words does not fit", (unsigned long long)cap, (unsigned long long)st->stride);
        return STORM_B200_ENOMEM;
    }
    STORM_CUDA_TRY(cudaMemsetAsync(fresh, 0, cap * st->stride * sizeof(uint64_t), st->stream));
    if (st->d_rows && st->uploaded_rows)
        STORM_CUDA_TRY(cudaMemcpyAsync(fresh, st->d_rows, st->uploaded_rows * st->stride * sizeof(uint64_t),
                                       cudaMemcpyDeviceToDevice, st->stream));
    if (st->d_rows) {
        STORM_CUDA_TRY(cudaStreamSynchronize(st->stream));
        cudaFree(st->d_rows);
    }
    st->d_rows = fresh;
    st->d_cap_rows = cap;
    return STORM_B200_OK;
}

// Bring rows [uploaded_rows, n_data) of the host mirror to the device.
int sync_rows(STORM_contiguous_t* c, ContigState* st) {
    int rc = ensure_device_rows(c, st, std::max<uint64_t>(c->n_data, 1));
    if (rc) return rc;
    if (st->uploaded_rows < c->n_data) {
        const uint64_t W = c->n_bitmaps_vector, r0 = st->uploaded_rows, n = c->n_data - r0;
        STORM_CUDA_TRY(cudaMemcpy2DAsync(st->d_rows + r0 * st->stride, st->stride * 8, c->data + r0 * W, W * 8,
                                         W * 8, n, cudaMemcpyHostToDevice, st->stream));
        st->uploaded_rows = c->n_data;
    }
    return STORM_B200_OK;
}

template <typename T>
int grow_device(T** p, uint64_t* cap, uint64_t need) {
    if (need <= *cap) return STORM_B200_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    uint64_t n = std::max<uint64_t>(need, *cap * 2);
    if (cudaMalloc(p, n * sizeof(T)) != cudaSuccess) { cudaGetLastError(); *cap = 0; set_error("device allocation of %llu bytes failed", (unsigned long long)(n * sizeof(T))); return STORM_B200_ENOMEM; }
    *cap = n;
    return STORM_B200_OK;
}

// Upload the sparse-row metadata used by the *_list entry points.
int sync_lists(STORM_contiguous_t* c, ContigState* st) {
    if (st->list_rows_synced == c->n_data && st->d_pos_off) return STORM_B200_OK;
    const uint64_t n = c->n_data;
    std::vector<uint64_t> off(n + 1);
    std::vector<uint32_t> is_sparse(n), sparse_rows, dense_rows;
    for (uint64_t r = 0; r < n; ++r) {
        const bool sp = c->n_scalar[r] < c->scalar_cutoff;
        is_sparse[r] = sp;
        off[r] = st->pos_off[r];
        (sp ? sparse_rows : dense_rows).push_back((uint32_t)r);
    }
    off[n] = c->tot_scalar;
    // off[r+1]-off[r] must be the list length for sparse rows: dense rows store nothing,
    // so consecutive offsets already delimit each sparse row's list.
    if (st->d_meta_cap < n + 1) {
        for (void* p : {(void*)st->d_pos_off, (void*)st->d_is_sparse, (void*)st->d_sparse_rows, (void*)st->d_dense_rows, (void*)st->d_group_start})
            if (p) cudaFree(p);
        const uint64_t cap = (n + 1) * 2;
        STORM_CUDA_TRY(cudaMalloc(&st->d_group_start, (cap + 1) * sizeof(uint32_t)));
        STORM_CUDA_TRY(cudaMalloc(&st->d_pos_off, cap * sizeof(uint64_t)));
        STORM_CUDA_TRY(cudaMalloc(&st->d_is_sparse, cap * sizeof(uint32_t)));
        STORM_CUDA_TRY(cudaMalloc(&st->d_sparse_rows, cap * sizeof(uint32_t)));
        STORM_CUDA_TRY(cudaMalloc(&st->d_dense_rows, cap * sizeof(uint32_t)));
        st->d_meta_cap = cap;
    }
    int rc = grow_device(&st->d_pos, &st->d_pos_cap, std::max<uint64_t>(c->tot_scalar, 1));
    if (rc) return rc;
    STORM_CUDA_TRY(cudaMemcpy(st->d_pos_off, off.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
    STORM_CUDA_TRY(cudaMemcpy(st->d_is_sparse, is_sparse.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (!sparse_rows.empty())
        STORM_CUDA_TRY(cudaMemcpy(st->d_sparse_rows, sparse_rows.data(), sparse_rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (!dense_rows.empty())
        STORM_CUDA_TRY(cudaMemcpy(st->d_dense_rows, dense_rows.data(), dense_rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (c->tot_scalar)
        STORM_CUDA_TRY(cudaMemcpy(st->d_pos, c->scalar, c->tot_scalar * sizeof(uint32_t), cudaMemcpyHostToDevice));
    st->n_sparse = sparse_rows.size();
    st->n_dense = dense_rows.size();
    // every row sparse: the lists are a complete flat form of the matrix, which the row-group stream kernel can answer from
    st->stream_ok = dense_rows.empty() && stream_groups(c->n_scalar, n, &st->h_group_start);
    if (st->stream_ok)
        STORM_CUDA_TRY(cudaMemcpy(st->d_group_start, st->h_group_start.data(), st->h_group_start.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    st->list_rows_synced = n;
    return STORM_B200_OK;
}

// ---- probe kernel: storm.c:108-129 on the device ---------------------------------
//
// One CTA per (sparse row s, slice of partner rows).  The CTA keeps s's positions
// (< scalar_cutoff <= 200 of them) in shared memory; each warp takes partner rows x
// and its lanes probe the positions into x's bitmap.  A pair {s, x} is counted
// once: x ranges over dense rows and over sparse rows with a larger row index.
constexpr int PROBE_THREADS = 256;
constexpr int PROBE_SLICE = 2048;          // partner rows per CTA
constexpr int PROBE_MAX_POS = 256;

__global__ void __launch_bounds__(PROBE_THREADS) contig_probe_kernel(
    const uint64_t* __restrict__ rows, uint64_t stride, uint64_t n_rows,
    const uint32_t* __restrict__ is_sparse, const uint32_t* __restrict__ sparse_rows,
    const uint32_t* __restrict__ pos, const uint64_t* __restrict__ pos_off,
    unsigned long long* total) {
    __shared__ uint32_t s_pos[PROBE_MAX_POS];
    __shared__ unsigned long long warp_part[PROBE_THREADS / 32];
    const uint32_t s = sparse_rows[blockIdx.x];
    const uint64_t b = pos_off[s];
    const uint32_t n = (uint32_t)(pos_off[s + 1] - b);
    for (uint32_t k = threadIdx.x; k < n; k += PROBE_THREADS) s_pos[k] = pos[b + k];
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t x0 = (uint64_t)blockIdx.y * PROBE_SLICE;
    const uint64_t x1 = min(x0 + (uint64_t)PROBE_SLICE, n_rows);
    unsigned long long cnt = 0;
    for (uint64_t x = x0 + warp; x < x1; x += PROBE_THREADS / 32) {
        if (x == s) continue;
        if (is_sparse[x] && x < s) continue;           // {x, s} is counted by x's CTA
        const uint64_t* row = rows + x * stride;
        for (uint32_t k = lane; k < n; k += 32) {
            const uint32_t p = s_pos[k];
            cnt += (row[p >> 6] >> (p & 63)) & 1ull;
        }
    }
    cnt = warp_sum(cnt);
    if (lane == 0) warp_part[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < PROBE_THREADS / 32; ++w) t += warp_part[w];
        if (t) atomicAdd(total, t);
    }
}

__global__ void gather_rows_kernel(uint64_t* dst, const uint64_t* src, uint64_t stride, const uint32_t* idx, uint64_t n) {
    // one CTA per destination row, 16-byte vectors
    const uint64_t r = blockIdx.x;
    if (r >= n) return;
    const uint4* s = reinterpret_cast<const uint4*>(src + (uint64_t)idx[r] * stride);
    uint4* d = reinterpret_cast<uint4*>(dst + r * stride);
    for (uint64_t k = threadIdx.x; k < stride / 2; k += blockDim.x) d[k] = s[k];
}

}  // namespace

int launch_contig_probe(const uint64_t* d_rows, uint64_t stride, uint64_t n_rows,
                        const uint32_t* d_is_sparse, const uint32_t* d_sparse_rows, uint64_t n_sparse,
                        const uint32_t* d_pos, const uint64_t* d_pos_off,
                        unsigned long long* d_total, cudaStream_t stream) {
    if (n_sparse == 0 || n_rows < 2) return STORM_B200_OK;
    const unsigned slices = (unsigned)((n_rows + PROBE_SLICE - 1) / PROBE_SLICE);
    uint64_t done = 0;
    while (done < n_sparse) {                     // grid.x <= 2^31-1, grid.y <= 65535
        const unsigned n = (unsigned)std::min<uint64_t>(n_sparse - done, 1u << 30);
        dim3 grid(n, slices);
        contig_probe_kernel<<<grid, PROBE_THREADS, 0, stream>>>(d_rows, stride, n_rows, d_is_sparse,
                                                               d_sparse_rows + done, d_pos, d_pos_off, d_total);
        STORM_CUDA_TRY(cudaGetLastError());
        count_launch();
        done += n;
    }
    return STORM_B200_OK;
}

int launch_gather_rows(uint64_t* dst, const uint64_t* src, uint64_t stride, const uint32_t* d_idx, uint64_t n,
                       cudaStream_t stream) {
    if (n == 0) return STORM_B200_OK;
    gather_rows_kernel<<<(unsigned)n, 128, 0, stream>>>(dst, src, stride, d_idx, n);
    STORM_CUDA_TRY(cudaGetLastError());
    count_launch();
    return STORM_B200_OK;
}

namespace {

enum QueryMode { QUERY_DENSE = 0, QUERY_LIST = 1 };
int g_list_route = 0;   // STORM_b200_set_contig_list_route: 0 cost model, 1 tile kernel, 2 tile + probe kernels, 3 stream kernel

// Shared body of all contiguous queries.
uint64_t contig_query(STORM_contiguous_t* c, QueryMode mode, uint32_t shard, uint32_t n_shards, int kernel) {
    if (c == nullptr) return (uint64_t)-1;                               // storm.c:1150,1176
    ContigState* st = state_of(c);
    if (mode == QUERY_LIST) {
        if (c->scalar == nullptr) return (uint64_t)-2;                   // storm.c:1245
        if (c->n_scalar == nullptr) return (uint64_t)-3;                 // storm.c:1246
    }
    if (c->n_data < 2) return 0;
    DeviceGuard guard(st->device);
    if (ensure_device_state(st)) return (uint64_t)-1;
    const auto t0 = std::chrono::steady_clock::now();
    cudaEventRecord(st->ev[0], st->stream);
    if (sync_rows(c, st)) return (uint64_t)-1;
    if (cudaMemsetAsync(st->d_total, 0, sizeof(unsigned long long), st->stream) != cudaSuccess) return (uint64_t)-1;
    cudaEventRecord(st->ev[1], st->stream);

    int rc = STORM_B200_OK;
    bool hybrid = false, stream = false;
    if (mode == QUERY_LIST) {
        if ((rc = sync_lists(c, st))) return (uint64_t)-1;
        // storm.c:1253-1258 switches per pair between the probe and the bitmap kernel; both give |i AND j|, so the
        // switch is a cost decision here (fitted: tile kernel 6e13 wp/s; probe kernel 0.08 ns + 1.3 ps per value
        // and pair, profiles/r01_configs_c1_c4_full.jsonl; stream kernel: stream_seconds): all rows through the
        // tile kernel, or dense x dense pairs through it and the rest probed, or -- when every row is a list --
        // the row-group stream kernel over the lists.
        const double N = (double)c->n_data, W = (double)c->n_bitmaps_vector, pairs = 0.5 * N * (N - 1.0);
        const double avg = (double)c->tot_scalar / std::max(1.0, (double)st->n_sparse);
        const double nd = (double)st->n_dense;
        const double tile_all_s = pairs * W / 6e13 + 3e-5;
        const double hybrid_s = 0.5 * nd * (nd - 1.0) * W / 6e13 + (nd >= 2 ? 3e-5 + nd * W * 8 / 2e12 : 0.0) +
                                (pairs - 0.5 * nd * (nd - 1.0)) * (0.08e-9 + 1.3e-12 * avg) + 1e-5;
        const double stream_s = st->stream_ok ? stream_seconds(pairs, avg) : 1e30;
        if (st->n_sparse > 0) {
            stream = stream_s < tile_all_s && stream_s < hybrid_s;
            hybrid = !stream && hybrid_s < tile_all_s;
            if (g_list_route == 1) { stream = false; hybrid = false; }
            if (g_list_route == 2) { stream = false; hybrid = true; }
            if (g_list_route == 3) { stream = st->stream_ok; hybrid = !stream; }
        }
        st->last_list_route = stream ? 3 : hybrid ? 2 : 1;
    }
    if (stream) {
        rc = launch_sparse_stream(st->d_pos_off, st->d_pos, st->h_group_start, st->d_group_start, st->d_pos_off, st->d_pos,
                                  c->tot_scalar, 0, c->n_data, 0, c->n_data, 1, shard, n_shards, st->d_total, st->stream);
    } else if (!hybrid) {
        rc = pairw_triangle(st->d_rows, c->n_data, c->n_bitmaps_vector, st->stride, shard, n_shards, kernel,
                            reinterpret_cast<uint64_t*>(st->d_total), st->stream);
    } else {
        // per-pair dispatch of storm.c:1253-1258: dense x dense pairs -> tile kernel on the
        // compacted dense rows; every pair with a sparse row -> probe kernel.
        if (st->n_dense >= 2) {
            if (st->d_gather_cap < st->n_dense * st->stride) {
                if (st->d_gather) cudaFree(st->d_gather);
                st->d_gather = nullptr; st->d_gather_cap = 0;
                if (cudaMalloc(&st->d_gather, st->n_dense * st->stride * 8) != cudaSuccess) {
                    cudaGetLastError(); set_error("device allocation for the dense-row gather failed"); return (uint64_t)-1;
                }
                st->d_gather_cap = st->n_dense * st->stride;
            }
            rc = launch_gather_rows(st->d_gather, st->d_rows, st->stride, st->d_dense_rows, st->n_dense, st->stream);
            if (!rc) rc = pairw_triangle(st->d_gather, st->n_dense, c->n_bitmaps_vector, st->stride, shard, n_shards,
                                         kernel, reinterpret_cast<uint64_t*>(st->d_total), st->stream);
        }
        if (!rc) {
            uint64_t b, e;
            shard_range(st->n_sparse, shard, n_shards, &b, &e);
            rc = launch_contig_probe(st->d_rows, st->stride, c->n_data, st->d_is_sparse, st->d_sparse_rows + b, e - b,
                                     st->d_pos, st->d_pos_off, st->d_total, st->stream);
        }
    }
    if (rc) return (uint64_t)-1;
    cudaEventRecord(st->ev[2], st->stream);
    if (cudaMemcpyAsync(st->h_total, st->d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st->stream) != cudaSuccess ||
        cudaStreamSynchronize(st->stream) != cudaSuccess) {
        set_error("query failed: %s", cudaGetErrorString(cudaGetLastError()));
        return (uint64_t)-1;
    }
    float up = 0, kn = 0;
    cudaEventElapsedTime(&up, st->ev[0], st->ev[1]);
    cudaEventElapsedTime(&kn, st->ev[1], st->ev[2]);
    st->timing[0] = up * 1e-3;
    st->timing[1] = kn * 1e-3;
    st->timing[2] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return *st->h_total;
}

int grow_host_rows(STORM_contiguous_t* c, uint64_t need) {
    if (need <= c->m_data) return 0;
    uint64_t cap = std::max<uint64_t>({need, (uint64_t)512, c->m_data * 2});     // geometric, not +512 (storm.c:1078-1100)
    const uint64_t W = c->n_bitmaps_vector;
    void* fresh = nullptr;
    if (posix_memalign(&fresh, 64, std::max<uint64_t>(cap * W * 8, 64))) return -1;
    if (c->data) memcpy(fresh, c->data, c->n_data * W * 8);
    memset((uint64_t*)fresh + c->n_data * W, 0, (cap - c->n_data) * W * 8);
    uint32_t* ns = (uint32_t*)realloc(c->n_scalar, cap * sizeof(uint32_t));
    STORM_contiguous_bitmap_t* bm = (STORM_contiguous_bitmap_t*)realloc(c->bitmaps, cap * sizeof(STORM_contiguous_bitmap_t));
    if (!ns || !bm) { free(fresh); if (ns) c->n_scalar = ns; if (bm) c->bitmaps = bm; return -1; }
    free(c->data);
    c->data = (uint64_t*)fresh; c->n_scalar = ns; c->bitmaps = bm; c->m_data = cap;
    for (uint64_t i = 0; i < cap; ++i) {
        c->bitmaps[i].data = c->data + i * W;
        if (i >= c->n_data) { c->bitmaps[i].scalar = nullptr; c->bitmaps[i].n_scalar = 0; }
    }
    return 0;
}

int grow_host_scalar(STORM_contiguous_t* c, ContigState* st, uint64_t need) {
    if (need <= c->m_scalar) return 0;
    uint64_t cap = std::max<uint64_t>({need, (uint64_t)512 * 32, c->m_scalar * 2});
    uint32_t* fresh = (uint32_t*)realloc(c->scalar, cap * sizeof(uint32_t));
    if (!fresh) return -1;
    c->scalar = fresh; c->m_scalar = cap;
    for (uint64_t i = 0; i < c->n_data; ++i)                               // re-point the per-row views (done right, D2)
        c->bitmaps[i].scalar = c->n_scalar[i] < c->scalar_cutoff ? c->scalar + st->pos_off[i] : nullptr;
    return 0;
}

// ---- scratch arena for the raw-buffer wrappers ----------------------------------
constexpr int MAX_UPLOAD_CHUNKS = 16;
constexpr uint64_t MIN_UPLOAD_CHUNK_BYTES = 32ull << 20;
struct Scratch {
    std::mutex mu;
    int device = -1;
    uint64_t* d_rows = nullptr; uint64_t cap_words = 0;
    unsigned long long* d_total = nullptr; unsigned long long* h_total = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;          // uploads of a streamed query run ahead of its kernels here
    cudaEvent_t chunk_ready[MAX_UPLOAD_CHUNKS] = {};
    cudaEvent_t idle = nullptr;
};
Scratch g_scratch;

int scratch_prepare(uint64_t words) {
    int rc = require_device();
    if (rc) return rc;
    int dev = 0;
    STORM_CUDA_TRY(cudaGetDevice(&dev));
    Scratch& s = g_scratch;
    if (s.device != dev) {                       // first use, or the caller switched device
        if (s.d_rows) cudaFree(s.d_rows);
        if (s.d_total) cudaFree(s.d_total);
        if (s.h_total) cudaFreeHost(s.h_total);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.copy_stream) cudaStreamDestroy(s.copy_stream);
        for (cudaEvent_t& e : s.chunk_ready) { if (e) cudaEventDestroy(e); e = nullptr; }
        if (s.idle) cudaEventDestroy(s.idle);
        s.d_rows = nullptr; s.cap_words = 0; s.d_total = nullptr; s.h_total = nullptr; s.stream = nullptr;
        s.copy_stream = nullptr; s.idle = nullptr;
        s.device = dev;
        STORM_CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        STORM_CUDA_TRY(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
        for (cudaEvent_t& e : s.chunk_ready) STORM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        STORM_CUDA_TRY(cudaEventCreateWithFlags(&s.idle, cudaEventDisableTiming));
        STORM_CUDA_TRY(cudaMalloc(&s.d_total, 8));
        STORM_CUDA_TRY(cudaMallocHost(&s.h_total, 8));
    }
    if (words > s.cap_words) {
        if (s.d_rows) cudaFree(s.d_rows);
        s.d_rows = nullptr; s.cap_words = 0;
        if (cudaMalloc(&s.d_rows, words * 8) != cudaSuccess) {
            cudaGetLastError(); set_error("scratch arena of %llu bytes does not fit", (unsigned long long)(words * 8));
            return STORM_B200_ENOMEM;
        }
        s.cap_words = words;
    }
    return STORM_B200_OK;
}

inline uint64_t padded_stride(uint64_t n_words) { return (n_words + ROW_ALIGN_WORDS - 1) / ROW_ALIGN_WORDS * ROW_ALIGN_WORDS; }

// Upload a host matrix with row pitch n_ints into the scratch arena at word offset `at`.
int scratch_upload(const uint64_t* vals, uint64_t n_vectors, uint64_t n_ints, uint64_t stride, uint64_t at) {
    Scratch& s = g_scratch;
    if (stride != n_ints)    // padding columns must read as zero
        STORM_CUDA_TRY(cudaMemsetAsync(s.d_rows + at, 0, n_vectors * stride * 8, s.stream));
    STORM_CUDA_TRY(cudaMemcpy2DAsync(s.d_rows + at, stride * 8, vals, n_ints * 8, n_ints * 8, n_vectors,
                                     cudaMemcpyHostToDevice, s.stream));
    return STORM_B200_OK;
}

// Host-buffer query with the upload hidden behind the kernels.  The triangle raster is monotone in the
// largest row a tile touches (common.cuh), so the matrix is uploaded in row chunks on a copy stream and
// the tiles that only need the rows uploaded so far are launched as soon as their chunk has landed:
// the query takes one chunk of PCIe time plus the kernel time instead of the whole upload plus the kernels.
int streamed_triangle(const uint64_t* vals, uint64_t n_vectors, uint64_t n_ints, uint64_t stride,
                      uint32_t shard, uint32_t n_shards, int kernel) {
    Scratch& s = g_scratch;
    kernel = resolve_kernel_for_rows(kernel, s.d_rows, n_vectors, (uint32_t)n_ints, stride);
    const TileShape ts = tile_shape_for(kernel);
    std::vector<uint64_t> prefix;
    uint32_t nbi = 0, nbj = 0;
    const uint64_t n_tiles = triangle_prefix(n_vectors, ts, &prefix, &nbi, &nbj);
    uint64_t tb = 0, te = 0;
    shard_range(n_tiles, shard, n_shards, &tb, &te);
    const uint64_t group_rows = (uint64_t)TRI_GROUP * ts.tn;               // rows a raster group adds
    const uint64_t n_groups = prefix.size() - 1;
    // chunks of whole groups, at least MIN_UPLOAD_CHUNK_BYTES each, at most MAX_UPLOAD_CHUNKS of them
    uint64_t groups_per_chunk = std::max<uint64_t>(1, (MIN_UPLOAD_CHUNK_BYTES + group_rows * n_ints * 8 - 1) / (group_rows * n_ints * 8));
    groups_per_chunk = std::max(groups_per_chunk, (n_groups + MAX_UPLOAD_CHUNKS - 1) / MAX_UPLOAD_CHUNKS);
    // the copy stream must not overwrite the arena while an earlier query's kernels still read it
    STORM_CUDA_TRY(cudaEventRecord(s.idle, s.stream));
    STORM_CUDA_TRY(cudaStreamWaitEvent(s.copy_stream, s.idle, 0));
    int chunk = 0;
    for (uint64_t g0 = 0; g0 < n_groups; g0 += groups_per_chunk, ++chunk) {
        const uint64_t g1 = std::min(n_groups, g0 + groups_per_chunk);
        const uint64_t r0 = g0 * group_rows, r1 = std::min<uint64_t>(n_vectors, g1 * group_rows);
        if (r1 > r0) {
            uint64_t* dst = s.d_rows + r0 * stride;
            if (stride != n_ints) STORM_CUDA_TRY(cudaMemsetAsync(dst, 0, (r1 - r0) * stride * 8, s.copy_stream));
            STORM_CUDA_TRY(cudaMemcpy2DAsync(dst, stride * 8, vals + r0 * n_ints, n_ints * 8, n_ints * 8, r1 - r0,
                                             cudaMemcpyHostToDevice, s.copy_stream));
        }
        STORM_CUDA_TRY(cudaEventRecord(s.chunk_ready[chunk], s.copy_stream));
        const uint64_t t0 = std::max(tb, prefix[g0]), t1 = std::min(te, prefix[g1]);
        if (t1 > t0) {
            STORM_CUDA_TRY(cudaStreamWaitEvent(s.stream, s.chunk_ready[chunk], 0));
            int rc = pairw_triangle_range(s.d_rows, n_vectors, (uint32_t)n_ints, stride, t0, t1, kernel,
                                          reinterpret_cast<uint64_t*>(s.d_total), s.stream);
            if (rc) return rc;
        }
    }
    // later queries on s.stream may reuse the arena: order them after the last upload as well
    STORM_CUDA_TRY(cudaStreamWaitEvent(s.stream, s.chunk_ready[chunk - 1], 0));
    return STORM_B200_OK;
}

uint64_t wrapper_diag_impl(uint64_t n_vectors, const uint64_t* vals, uint64_t n_ints,
                           uint32_t shard = 0, uint32_t n_shards = 1, int kernel = STORM_B200_KERNEL_AUTO,
                           int op = STORM_B200_OP_INTERSECT) {
    if (vals == nullptr || n_ints == 0) { set_error("STORM_wrapper_diag: NULL buffer or zero width"); return (uint64_t)-1; }
    if (n_shards == 0 || shard >= n_shards) { set_error("shard %u of %u", shard, n_shards); return (uint64_t)-1; }
    if (n_vectors < 2) return 0;
    std::lock_guard<std::mutex> lock(g_scratch.mu);
    const uint64_t stride = padded_stride(n_ints);
    if (scratch_prepare(n_vectors * stride)) return (uint64_t)-1;
    Scratch& s = g_scratch;
    if (cudaMemsetAsync(s.d_total, 0, 8, s.stream) != cudaSuccess) return (uint64_t)-1;
    if (op == STORM_B200_OP_INTERSECT && n_vectors * n_ints * 8 >= 2 * MIN_UPLOAD_CHUNK_BYTES) {
        if (streamed_triangle(vals, n_vectors, n_ints, stride, shard, n_shards, kernel)) return (uint64_t)-1;
    } else if (scratch_upload(vals, n_vectors, n_ints, stride, 0)) {
        return (uint64_t)-1;
    } else if (op != STORM_B200_OP_INTERSECT) {
        if (n_shards != 1) { set_error("set operations other than intersect are not sharded"); return (uint64_t)-1; }
        if (pairw_rect_op(s.d_rows, n_vectors, stride, 0, s.d_rows, n_vectors, stride, 0, (uint32_t)n_ints, 1, op, kernel, true,
                          nullptr, 0, reinterpret_cast<uint64_t*>(s.d_total), s.stream)) return (uint64_t)-1;
    } else if (pairw_triangle(s.d_rows, n_vectors, (uint32_t)n_ints, stride, shard, n_shards, kernel,
                              reinterpret_cast<uint64_t*>(s.d_total), s.stream)) return (uint64_t)-1;
    if (cudaMemcpyAsync(s.h_total, s.d_total, 8, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess ||
        cudaStreamSynchronize(s.stream) != cudaSuccess) {
        set_error("STORM_wrapper_diag failed: %s", cudaGetErrorString(cudaGetLastError()));
        return (uint64_t)-1;
    }
    return *s.h_total;
}

}  // namespace
}  // namespace storm

// =================================================================================
// C ABI: storm.h contiguous entry points
// =================================================================================
extern "C" {
uint64_t STORM_b200_host_union_count(const uint64_t* a, const uint64_t* b, const size_t n);
uint64_t STORM_b200_host_diff_count(const uint64_t* a, const uint64_t* b, const size_t n);
}
namespace storm {
int op_of_compute_func(const STORM_compute_func f) {
    if (f == &STORM_b200_host_union_count) return STORM_B200_OP_UNION;
    if (f == &STORM_b200_host_diff_count) return STORM_B200_OP_DIFF;
    return STORM_B200_OP_INTERSECT;
}
}  // namespace storm

using namespace storm;

extern "C" {

STORM_contiguous_t* STORM_contig_new(size_t vector_length) {               // storm.c:1001-1018
    STORM_contiguous_t* c = (STORM_contiguous_t*)calloc(1, sizeof(STORM_contiguous_t));
    if (c == nullptr) return nullptr;
    ContigState* st = new (std::nothrow) ContigState();
    if (st == nullptr) { free(c); return nullptr; }
    c->b200 = st;
    c->vector_length = vector_length;
    c->n_bitmaps_vector = (uint32_t)((vector_length + 63) / 64);
    c->alignment = 64;
    c->intsec_func = &host_intersect_count;
    c->scalar_cutoff = (uint32_t)(vector_length / 200 > 200 ? 200 : vector_length / 200);
    int n = 0;                                   // remember the creating thread's device, if any
    if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) cudaGetDevice(&st->device); else cudaGetLastError();
    return c;
}

void STORM_contig_free(STORM_contiguous_t* c) {                             // storm.c:1020-1029
    if (c == nullptr) return;
    ContigState* st = state_of(c);
    if (st) {
        DeviceGuard guard(st->device);
        if (st->stream) cudaStreamSynchronize(st->stream);
        for (void* p : {(void*)st->d_rows, (void*)st->d_pos, (void*)st->d_pos_off, (void*)st->d_is_sparse,
                        (void*)st->d_sparse_rows, (void*)st->d_dense_rows, (void*)st->d_group_start, (void*)st->d_gather, (void*)st->d_total})
            if (p) cudaFree(p);
        if (st->h_total) cudaFreeHost(st->h_total);
        for (auto e : st->ev) if (e) cudaEventDestroy(e);
        if (st->stream) cudaStreamDestroy(st->stream);
        delete st;
    }
    free(c->data); free(c->scalar); free(c->n_scalar); free(c->bitmaps);
    free(c);
}

int STORM_contig_add(STORM_contiguous_t* c, const uint32_t* values, const uint32_t n_values) {   // storm.c:1031-1137
    if (c == nullptr) return -1;
    if (values == nullptr) return -2;
    if (n_values == 0) return 0;                                          // no row appended (D7)
    ContigState* st = state_of(c);
    for (uint32_t i = 0; i < n_values; ++i)
        if (values[i] >= c->vector_length) { set_error("position %u >= vector_length %llu", values[i], (unsigned long long)c->vector_length); return -3; }
    if (grow_host_rows(c, c->n_data + 1)) return -3;
    if (c->scalar == nullptr && grow_host_scalar(c, st, 1)) return -3;    // reference allocates it on first add (:1037-1041)

    uint64_t* row = c->data + c->n_data * c->n_bitmaps_vector;
    uint32_t used = n_values;
    for (uint32_t i = 0; i < n_values; ++i) {                             // :1103-1115
        if (i != 0 && values[i] == values[i - 1]) { --used; continue; }
        row[values[i] >> 6] |= 1ull << (values[i] & 63);
    }
    if (st->pos_off.size() <= c->n_data) st->pos_off.resize(std::max<size_t>(c->n_data + 1, st->pos_off.size() * 2));
    st->pos_off[c->n_data] = c->tot_scalar;
    STORM_contiguous_bitmap_t* view = &c->bitmaps[c->n_data];
    view->scalar = nullptr;
    if (used < c->scalar_cutoff) {                                        // :1119-1129, compacted (D11)
        if (grow_host_scalar(c, st, c->tot_scalar + used)) return -3;
        uint32_t* dst = c->scalar + c->tot_scalar;
        uint32_t w = 0;
        for (uint32_t i = 0; i < n_values; ++i) {
            if (i != 0 && values[i] == values[i - 1]) continue;
            dst[w++] = values[i];
        }
        view->scalar = dst;
        c->tot_scalar += used;
    }
    c->n_scalar[c->n_data] = used;                                        // :1132-1133
    view->n_scalar = used;
    ++c->n_data;                                                          // :1134
    return (int)n_values;                                                 // :1136
}

int STORM_contig_clear(STORM_contiguous_t* c) {                            // storm.c:1139-1147
    if (c == nullptr) return -1;
    if (c->data == nullptr) return 0;
    memset(c->data, 0, (uint64_t)c->n_bitmaps_vector * c->m_data * sizeof(uint64_t));
    c->n_data = 0;
    c->tot_scalar = 0;
    ContigState* st = state_of(c);
    st->uploaded_rows = 0;
    st->list_rows_synced = 0;
    if (st->d_rows) {
        DeviceGuard guard(st->device);
        cudaMemsetAsync(st->d_rows, 0, st->d_cap_rows * st->stride * 8, st->stream);
    }
    return 1;
}

uint64_t STORM_contig_pairw_intersect_cardinality(STORM_contiguous_t* c) {          // storm.c:1149-1173
    if (c == nullptr) return (uint64_t)-1;
    return contig_query(c, QUERY_DENSE, 0, 1, STORM_B200_KERNEL_AUTO);
}

uint64_t STORM_contig_pairw_intersect_cardinality_blocked(STORM_contiguous_t* c, uint32_t bsize) {   // :1175-1241
    (void)bsize;   // CPU cache-blocking hint; the tile shape is fixed by the kernel
    if (c == nullptr) return (uint64_t)-1;
    return contig_query(c, QUERY_DENSE, 0, 1, STORM_B200_KERNEL_AUTO);
}

int STORM_b200_set_contig_list_route(int route) {
    const int prev = g_list_route;
    if (route >= 0 && route <= 3) g_list_route = route;
    return prev;
}

int STORM_b200_contig_last_list_route(const STORM_contiguous_t* c) {
    if (c == nullptr || c->b200 == nullptr) return 0;
    return state_of(const_cast<STORM_contiguous_t*>(c))->last_list_route;
}

uint64_t STORM_contig_pairw_intersect_cardinality_list(STORM_contiguous_t* c) {     // storm.c:1243-1263
    if (c == nullptr) return (uint64_t)-1;
    return contig_query(c, QUERY_LIST, 0, 1, STORM_B200_KERNEL_AUTO);
}

uint64_t STORM_contig_pairw_intersect_cardinality_blocked_list(STORM_contiguous_t* c, uint32_t bsize) {  // :1265-1347
    (void)bsize;
    if (c == nullptr) return (uint64_t)-1;
    return contig_query(c, QUERY_LIST, 0, 1, STORM_B200_KERNEL_AUTO);
}

// ---- raw-buffer wrappers (storm.c:132-369) ---------------------------------------
// ---- per-pair kernel pointers (libalgebra.h:3035, 3094-3236) ----------------------------------
// The reference's raw-buffer loops apply whatever STORM_compute_func the caller hands them.  Host code
// cannot run on the device, but the three families libalgebra offers can be told apart by address:
// these exported functions are what the STORM_get_*_count_func choosers of this library return, and a
// wrapper that receives the union or diff one answers with that set operation (setops.cu).  Any other
// pointer, including NULL and the reference's own static kernels, means intersect.
uint64_t STORM_b200_host_intersect_count(const uint64_t* a, const uint64_t* b, const size_t n) {
    return host_intersect_count(a, b, n);
}
uint64_t STORM_b200_host_union_count(const uint64_t* a, const uint64_t* b, const size_t n) {       // libalgebra.h:2994-3000
    uint64_t c = 0;
    for (size_t k = 0; k < n; ++k) c += (uint64_t)__builtin_popcountll(a[k] | b[k]);
    return c;
}
uint64_t STORM_b200_host_diff_count(const uint64_t* a, const uint64_t* b, const size_t n) {        // libalgebra.h:3002-3008
    uint64_t c = 0;
    for (size_t k = 0; k < n; ++k) c += (uint64_t)__builtin_popcountll(a[k] ^ b[k]);
    return c;
}
STORM_compute_func STORM_get_intersect_count_func(const size_t n_bitmaps_vector) { (void)n_bitmaps_vector; return &STORM_b200_host_intersect_count; }
STORM_compute_func STORM_get_union_count_func(const size_t n_bitmaps_vector) { (void)n_bitmaps_vector; return &STORM_b200_host_union_count; }
STORM_compute_func STORM_get_diff_count_func(const size_t n_bitmaps_vector) { (void)n_bitmaps_vector; return &STORM_b200_host_diff_count; }

uint64_t STORM_wrapper_diag(const uint32_t n_vectors, const uint64_t* vals, const uint32_t n_ints, const STORM_compute_func f) {
    return wrapper_diag_impl(n_vectors, vals, n_ints, 0, 1, STORM_B200_KERNEL_AUTO, op_of_compute_func(f));
}

uint64_t STORM_wrapper_diag_blocked(const uint32_t n_vectors, const uint64_t* vals, const uint32_t n_ints,
                                    const STORM_compute_func f, uint32_t block_size) {
    (void)block_size;
    return wrapper_diag_impl(n_vectors, vals, n_ints, 0, 1, STORM_B200_KERNEL_AUTO, op_of_compute_func(f));
}

uint64_t STORM_wrapper_square(const uint32_t n_vectors1, const uint64_t* STORM_RESTRICT vals1,
                              const uint32_t n_vectors2, const uint64_t* STORM_RESTRICT vals2,
                              const uint32_t n_ints, const STORM_compute_func f) {
    const int op = op_of_compute_func(f);
    if (!vals1 || !vals2 || n_ints == 0) { set_error("STORM_wrapper_square: NULL buffer or zero width"); return (uint64_t)-1; }
    if (n_vectors1 == 0 || n_vectors2 == 0) return 0;
    std::lock_guard<std::mutex> lock(g_scratch.mu);
    const uint64_t stride = padded_stride(n_ints);
    if (scratch_prepare(((uint64_t)n_vectors1 + n_vectors2) * stride)) return (uint64_t)-1;
    Scratch& s = g_scratch;
    const uint64_t at2 = (uint64_t)n_vectors1 * stride;
    if (scratch_upload(vals1, n_vectors1, n_ints, stride, 0) || scratch_upload(vals2, n_vectors2, n_ints, stride, at2)) return (uint64_t)-1;
    if (cudaMemsetAsync(s.d_total, 0, 8, s.stream) != cudaSuccess) return (uint64_t)-1;
    if (pairw_rect_op(s.d_rows, n_vectors1, stride, 0, s.d_rows + at2, n_vectors2, stride, 0, n_ints, 0, op,
                      STORM_B200_KERNEL_AUTO, false, nullptr, 0, reinterpret_cast<uint64_t*>(s.d_total), s.stream)) return (uint64_t)-1;
    if (cudaMemcpyAsync(s.h_total, s.d_total, 8, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess ||
        cudaStreamSynchronize(s.stream) != cudaSuccess) {
        set_error("STORM_wrapper_square failed: %s", cudaGetErrorString(cudaGetLastError()));
        return (uint64_t)-1;
    }
    return *s.h_total;
}

// The list wrappers take caller-built position arrays (storm.c:173-220, 281-369).
// They are answered from the bitmaps alone: the value is identical (the lists only
// select a cheaper CPU code path per pair in the reference).
uint64_t STORM_wrapper_diag_list(const uint32_t n_vectors, const uint64_t* STORM_RESTRICT vals, const uint32_t n_ints,
                                 const uint32_t* STORM_RESTRICT n_alts, const uint32_t* STORM_RESTRICT alt_positions,
                                 const uint32_t* STORM_RESTRICT alt_offsets, const STORM_compute_func f,
                                 const STORM_compute_lfunc fl, const uint32_t cutoff) {
    (void)n_alts; (void)alt_positions; (void)alt_offsets; (void)f; (void)fl; (void)cutoff;
    return wrapper_diag_impl(n_vectors, vals, n_ints);
}

uint64_t STORM_wrapper_diag_list_blocked(const uint32_t n_vectors, const uint64_t* STORM_RESTRICT vals, const uint32_t n_ints,
                                         const uint32_t* STORM_RESTRICT n_alts, const uint32_t* STORM_RESTRICT alt_positions,
                                         const uint32_t* STORM_RESTRICT alt_offsets, const STORM_compute_func f,
                                         const STORM_compute_lfunc fl, const uint32_t cutoff, uint32_t block_size) {
    (void)n_alts; (void)alt_positions; (void)alt_offsets; (void)f; (void)fl; (void)cutoff; (void)block_size;
    return wrapper_diag_impl(n_vectors, vals, n_ints);
}

uint64_t STORM_b200_wrapper_diag_shard(uint64_t n_vectors, const uint64_t* vals, uint64_t n_ints,
                                       uint32_t shard, uint32_t n_shards, int kernel) {
    return wrapper_diag_impl(n_vectors, vals, n_ints, shard, n_shards, kernel);
}

// ---- storm_b200.h container extensions -------------------------------------------
uint64_t STORM_b200_contig_pairw_shard(STORM_contiguous_t* c, uint32_t shard, uint32_t n_shards, int kernel) {
    if (c == nullptr) return (uint64_t)-1;
    if (n_shards == 0 || shard >= n_shards) { set_error("shard %u of %u", shard, n_shards); return (uint64_t)-1; }
    return contig_query(c, QUERY_DENSE, shard, n_shards, kernel);
}

int STORM_b200_contig_pairw_rect(STORM_contiguous_t* c, uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, uint32_t* out) {
    if (c == nullptr || out == nullptr) { set_error("NULL argument"); return STORM_B200_EINVAL; }
    if (i0 > i1 || j0 > j1 || i1 > c->n_data || j1 > c->n_data) { set_error("rectangle outside the %llu rows", (unsigned long long)c->n_data); return STORM_B200_EINVAL; }
    if (i0 == i1 || j0 == j1) return STORM_B200_OK;
    ContigState* st = state_of(c);
    DeviceGuard guard(st->device);
    int rc = ensure_device_state(st);
    if (rc) return rc;
    if ((rc = sync_rows(c, st))) return rc;
    const uint64_t ni = i1 - i0, nj = j1 - j0;
    uint32_t* d_out = nullptr;
    if (cudaMalloc(&d_out, ni * nj * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); set_error("device allocation for %llu x %llu counts failed", (unsigned long long)ni, (unsigned long long)nj); return STORM_B200_ENOMEM; }
    rc = pairw_rect(st->d_rows + i0 * st->stride, ni, st->stride, i0, st->d_rows + j0 * st->stride, nj, st->stride, j0,
                    c->n_bitmaps_vector, 1, STORM_B200_KERNEL_AUTO, d_out, nj, nullptr, st->stream);
    if (!rc && (cudaMemcpyAsync(out, d_out, ni * nj * sizeof(uint32_t), cudaMemcpyDeviceToHost, st->stream) != cudaSuccess ||
                cudaStreamSynchronize(st->stream) != cudaSuccess)) {
        set_error("rect query failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = STORM_B200_ECUDA;
    }
    cudaFree(d_out);
    return rc;
}

const uint64_t* STORM_b200_contig_device_rows(STORM_contiguous_t* c, uint64_t* row_stride_words) {
    if (c == nullptr) return nullptr;
    ContigState* st = state_of(c);
    DeviceGuard guard(st->device);
    if (ensure_device_state(st) || sync_rows(c, st)) return nullptr;
    if (cudaStreamSynchronize(st->stream) != cudaSuccess) return nullptr;
    if (row_stride_words) *row_stride_words = st->stride;
    return st->d_rows;
}

int STORM_b200_contig_invalidate_device(STORM_contiguous_t* c) {
    if (c == nullptr) return STORM_B200_EINVAL;
    ContigState* st = state_of(c);
    st->uploaded_rows = 0;
    st->list_rows_synced = 0;
    return STORM_B200_OK;
}

int STORM_b200_contig_add_bulk(STORM_contiguous_t* c, const uint32_t* positions, const uint64_t* offsets, uint64_t n_rows) {
    if (c == nullptr || offsets == nullptr || (positions == nullptr && n_rows && offsets[n_rows] != offsets[0])) { set_error("NULL argument"); return STORM_B200_EINVAL; }
    if (n_rows == 0) return STORM_B200_OK;
    ContigState* st = state_of(c);
    DeviceGuard guard(st->device);
    int rc = ensure_device_state(st);
    if (rc) return rc;
    // host pass: validate, compact away empty rows (D7) and adjacent duplicates, record list metadata.  The checks
    // are one tight loop per row (range, adjacent duplicates); a clean row is taken over with one memcpy, and when
    // EVERY row is clean and non-empty -- what callers that sort + unique first (benchmark.cpp:571-572) hand in --
    // nothing is copied at all: the device reads the caller's buffer.
    std::vector<uint64_t> off;
    std::vector<uint32_t> pos;
    off.reserve(n_rows + 1);
    off.push_back(0);
    std::vector<uint32_t> used_per_row;
    used_per_row.reserve(n_rows);
    bool all_clean = true;                      // no empty row, no duplicate so far: `pos` has not been started
    for (uint64_t r = 0; r < n_rows; ++r) {
        const uint64_t b = offsets[r], e = offsets[r + 1];
        if (e < b) { set_error("offsets must be non-decreasing"); return STORM_B200_EINVAL; }
        uint32_t mx = 0, dups = 0;
        const uint32_t* p = positions + b;
        const uint64_t n = e - b;
        for (uint64_t k = 0; k < n; ++k) mx = std::max(mx, p[k]);
        for (uint64_t k = 1; k < n; ++k) dups += p[k] == p[k - 1];
        if (n && (uint64_t)mx >= c->vector_length) { set_error("position %u >= vector_length", mx); return STORM_B200_EINVAL; }
        const bool clean = n > 0 && dups == 0;
        if (all_clean && !clean) {              // first row that needs compaction: materialise what was skipped so far
            all_clean = false;
            pos.reserve(offsets[n_rows] - offsets[0]);
            pos.assign(positions + offsets[0], positions + b);
        }
        if (n == 0) continue;
        if (!all_clean) {
            if (dups == 0) pos.insert(pos.end(), p, p + n);
            else
                for (uint64_t k = 0; k < n; ++k)
                    if (k == 0 || p[k] != p[k - 1]) pos.push_back(p[k]);
        }
        const uint64_t end = off.back() + (n - dups);
        used_per_row.push_back((uint32_t)(n - dups));
        off.push_back(end);
    }
    const uint32_t* src_pos = all_clean ? positions + offsets[0] : pos.data();     // compact, duplicate-free positions
    const uint64_t n_pos = off.back();
    const uint64_t n_new = used_per_row.size();
    if (n_new == 0) return STORM_B200_OK;
    if (grow_host_rows(c, c->n_data + n_new)) { set_error("host arena allocation failed"); return STORM_B200_ENOMEM; }
    if (c->scalar == nullptr && grow_host_scalar(c, st, 1)) return STORM_B200_ENOMEM;
    if ((rc = sync_rows(c, st))) return rc;                               // earlier rows first
    if ((rc = ensure_device_rows(c, st, c->n_data + n_new))) return rc;

    uint32_t* d_pos = nullptr; uint64_t* d_off = nullptr;
    STORM_CUDA_TRY(cudaMalloc(&d_pos, std::max<size_t>(n_pos, 1) * sizeof(uint32_t)));
    STORM_CUDA_TRY(cudaMalloc(&d_off, off.size() * sizeof(uint64_t)));
    STORM_CUDA_TRY(cudaMemcpyAsync(d_pos, src_pos, n_pos * sizeof(uint32_t), cudaMemcpyHostToDevice, st->stream));
    STORM_CUDA_TRY(cudaMemcpyAsync(d_off, off.data(), off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st->stream));
    uint64_t* d_dst = st->d_rows + c->n_data * st->stride;
    STORM_CUDA_TRY(cudaMemsetAsync(d_dst, 0, n_new * st->stride * 8, st->stream));
    rc = launch_scatter_positions(d_dst, st->stride, d_pos, d_off, n_new, st->stream);
    // keep the public host mirror valid: copy the scattered rows back
    const uint64_t W = c->n_bitmaps_vector;
    if (!rc && cudaMemcpy2DAsync(c->data + c->n_data * W, W * 8, d_dst, st->stride * 8, W * 8, n_new,
                                 cudaMemcpyDeviceToHost, st->stream) != cudaSuccess) rc = STORM_B200_ECUDA;
    if (cudaStreamSynchronize(st->stream) != cudaSuccess) rc = STORM_B200_ECUDA;
    cudaFree(d_pos); cudaFree(d_off);
    if (rc) { set_error("bulk ingest failed: %s", cudaGetErrorString(cudaGetLastError())); return rc; }

    if (st->pos_off.size() < c->n_data + n_new) st->pos_off.resize(c->n_data + n_new);
    for (uint64_t r = 0; r < n_new; ++r) {
        const uint64_t row = c->n_data + r;
        const uint32_t used = used_per_row[r];
        st->pos_off[row] = c->tot_scalar;
        c->bitmaps[row].scalar = nullptr;
        if (used < c->scalar_cutoff) {
            if (grow_host_scalar(c, st, c->tot_scalar + used)) return STORM_B200_ENOMEM;
            memcpy(c->scalar + c->tot_scalar, src_pos + off[r], used * sizeof(uint32_t));
            c->bitmaps[row].scalar = c->scalar + c->tot_scalar;
            c->tot_scalar += used;
        }
        c->n_scalar[row] = used;
        c->bitmaps[row].n_scalar = used;
    }
    c->n_data += n_new;
    st->uploaded_rows = c->n_data;            // the device already holds them
    return STORM_B200_OK;
}

int STORM_b200_contig_last_timing(STORM_contiguous_t* c, double out_seconds[3]) {
    if (c == nullptr || out_seconds == nullptr) return STORM_B200_EINVAL;
    ContigState* st = state_of(c);
    for (int i = 0; i < 3; ++i) out_seconds[i] = st->timing[i];
    return STORM_B200_OK;
}

}  // extern "C"
