// contig.cu -- STORM_contiguous_t: host container (reference-compatible public
// fields), device-resident 128-byte-aligned row arena, and the pairwise queries.
//
// Reference being replaced: storm.c:1001-1347 (container + four query loops) and
// storm.c:132-279 (raw-buffer wrappers).  Host code only builds and uploads the
// rows; every query is answered by CUDA kernels.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"
#include "devices.h"
#include "runtime.h"

namespace storm {

namespace {

constexpr uint64_t ROW_ALIGN_WORDS = 16;   // device row stride is a multiple of 128 bytes
constexpr uint64_t BG_UPLOAD_BYTES = 8ull << 20;   // STORM_contig_add pushes finished rows to the devices in batches of this size

// One replica of the container on one device.
struct ContigDev {
    DevCtx ctx;
    uint64_t* d_rows = nullptr;        // row arena (stride words per row)
    uint64_t d_cap_rows = 0;
    // sparse-row position lists and row classes (the *_list queries)
    uint32_t* d_pos = nullptr; uint64_t d_pos_cap = 0;
    uint64_t* d_pos_off = nullptr; uint32_t* d_is_sparse = nullptr; uint32_t* d_sparse_rows = nullptr;
    uint32_t* d_dense_rows = nullptr; uint32_t* d_group_start = nullptr; uint64_t d_meta_cap = 0;
    uint64_t list_rows_synced = 0;     // list metadata of rows [0, list_rows_synced) is on this device
    uint64_t* d_gather = nullptr; uint64_t d_gather_cap = 0;   // compact arena of dense rows (list path)
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};           // timing (primary device)
};

struct ContigState {
    int home_device = -1;              // current device of the creating thread: the default device set
    int dev_status = 0;                // 0 = devices not set up yet, 1 = ready, -1 = setting them up failed
    std::vector<ContigDev*> devs;      // devs[0] is the primary (rectangles, bulk ingest, set operations)
    uint64_t stride = 0;               // words, the same on every device
    uint64_t uploaded_rows = 0;        // rows [0, uploaded_rows) of the host mirror have been requested on every device's copy stream
    bool data_pinned = false;          // c->data came from cudaHostAlloc (portable): uploads are direct DMA
    // sparse-row position lists: per-row offsets into c->scalar, and the host form of the list metadata
    std::vector<uint64_t> pos_off;
    std::vector<uint64_t> h_off;
    std::vector<uint32_t> h_is_sparse, h_sparse_rows, h_dense_rows, h_group_start;
    uint64_t list_rows_built = (uint64_t)-1;
    uint64_t n_sparse = 0, n_dense = 0;
    bool stream_ok = false;            // every row is sparse and short enough for the stream kernel
    int last_list_route = 0;           // 1 tile kernel, 2 tile + probe kernels, 3 stream kernel
    double timing[3] = {0, 0, 0};
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

bool have_device() {
    static int state = 0;              // 0 unknown, 1 yes, 2 no
    if (state == 0) {
        int n = 0;
        state = (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) ? 1 : 2;
        if (state == 2) cudaGetLastError();
    }
    return state == 1;
}

// Device replicas of a container: the device set in force at its first use (devices.h), fixed for its lifetime.
int ensure_devs(ContigState* st) {
    if (st->dev_status == 1) return STORM_B200_OK;
    DeviceGuard home(st->home_device);               // "current device" means the creating thread's
    std::vector<int> ids;
    int rc = query_devices(&ids);
    if (rc) return rc;
    if (st->home_device < 0) st->home_device = ids[0];
    enable_peers(ids);
    for (int id : ids) {
        ContigDev* d = new (std::nothrow) ContigDev();
        if (!d) { set_error("out of host memory"); return STORM_B200_ENOMEM; }
        st->devs.push_back(d);                       // (a half-built replica is torn down by STORM_contig_free)
        if ((rc = d->ctx.init(id))) { st->dev_status = -1; return rc; }
    }
    {
        DeviceGuard guard(st->devs[0]->ctx.device);
        for (auto& e : st->devs[0]->ev) STORM_CUDA_TRY(cudaEventCreate(&e));
    }
    st->dev_status = 1;
    return STORM_B200_OK;
}

// One-time work a device's first dense query would otherwise pay (the FP4 accumulation self-test behind KERNEL_AUTO,
// ~0.1 s per device and process): done when the background uploads start, i.e. while the caller is still adding rows.
void warm_up_devices(ContigState* st) {
    for (ContigDev* d : st->devs) { DeviceGuard guard(d->ctx.device); (void)fp4_selftest_ok(); }
}

// Device arena of one replica with room for `rows` rows (grows geometrically, like the host mirror).
int ensure_device_rows(STORM_contiguous_t* c, ContigState* st, ContigDev* d, uint64_t rows) {
    if (st->stride == 0)
        st->stride = ((uint64_t)c->n_bitmaps_vector + ROW_ALIGN_WORDS - 1) / ROW_ALIGN_WORDS * ROW_ALIGN_WORDS;
    if (rows <= d->d_cap_rows) return STORM_B200_OK;
    DeviceGuard guard(d->ctx.device);
    uint64_t cap = std::max<uint64_t>({rows, c->m_data, d->d_cap_rows + d->d_cap_rows / 2});
    cap = (cap + 511) / 512 * 512;
    uint64_t* fresh = nullptr;
    if (cudaMalloc(&fresh, cap * st->stride * sizeof(uint64_t)) != cudaSuccess) {
        cudaGetLastError();
        set_error("device arena of %llu rows x %llu words does not fit on device %d", (unsigned long long)cap,
                  (unsigned long long)st->stride, d->ctx.device);
        return STORM_B200_ENOMEM;
    }
    STORM_CUDA_TRY(cudaMemsetAsync(fresh, 0, cap * st->stride * sizeof(uint64_t), d->ctx.copy_stream));
    if (d->d_rows && st->uploaded_rows)
        STORM_CUDA_TRY(cudaMemcpyAsync(fresh, d->d_rows, st->uploaded_rows * st->stride * sizeof(uint64_t),
                                       cudaMemcpyDeviceToDevice, d->ctx.copy_stream));
    if (d->d_rows) {
        STORM_CUDA_TRY(cudaStreamSynchronize(d->ctx.copy_stream));
        STORM_CUDA_TRY(cudaStreamSynchronize(d->ctx.stream));
        cudaFree(d->d_rows);
    }
    d->d_rows = fresh;
    d->d_cap_rows = cap;
    return STORM_B200_OK;
}

inline HostRows mirror_of(const STORM_contiguous_t* c, const ContigState* st) {
    return HostRows{c->data, c->n_bitmaps_vector, st->data_pinned};
}

// Request rows [uploaded_rows, upto) of the host mirror on every replica (each over its own PCIe link, copy stream).
int upload_pending(STORM_contiguous_t* c, ContigState* st, uint64_t upto) {
    for (ContigDev* d : st->devs) {
        int rc = ensure_device_rows(c, st, d, std::max<uint64_t>(upto, 1));
        if (rc) return rc;
        if (st->uploaded_rows < upto && (rc = upload_rows(&d->ctx, d->d_rows, st->stride, mirror_of(c, st), c->n_bitmaps_vector, st->uploaded_rows, upto))) return rc;
    }
    if (st->uploaded_rows < upto) st->uploaded_rows = upto;
    return STORM_B200_OK;
}

// Kernels on `stream` launched after this see every upload requested so far on the replica's copy stream.
int order_after_uploads(ContigDev* d) {
    DeviceGuard guard(d->ctx.device);
    STORM_CUDA_TRY(cudaEventRecord(d->ctx.mark, d->ctx.copy_stream));
    STORM_CUDA_TRY(cudaStreamWaitEvent(d->ctx.stream, d->ctx.mark, 0));
    return STORM_B200_OK;
}

// Block the host until no copy of any replica reads the host mirror any more (before it is freed).
void wait_uploads(ContigState* st) {
    for (ContigDev* d : st->devs)
        if (d->ctx.copy_stream) { DeviceGuard guard(d->ctx.device); cudaStreamSynchronize(d->ctx.copy_stream); }
}

template <typename T>
int grow_device(T** p, uint64_t* cap, uint64_t need) {
    if (need <= *cap) return STORM_B200_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    const uint64_t n = std::max<uint64_t>(need, *cap * 2);
    *cap = 0;
    if (cudaMalloc(p, n * sizeof(T)) != cudaSuccess) { cudaGetLastError(); *p = nullptr; set_error("device allocation of %llu bytes failed", (unsigned long long)(n * sizeof(T))); return STORM_B200_ENOMEM; }
    *cap = n;
    return STORM_B200_OK;
}

// Host form of the sparse-row metadata used by the *_list entry points (rebuilt when rows were added).
void build_list_meta(STORM_contiguous_t* c, ContigState* st) {
    if (st->list_rows_built == c->n_data) return;
    const uint64_t n = c->n_data;
    st->h_off.assign(n + 1, 0);
    st->h_is_sparse.assign(n, 0);
    st->h_sparse_rows.clear(); st->h_dense_rows.clear();
    for (uint64_t r = 0; r < n; ++r) {
        const bool sp = c->n_scalar[r] < c->scalar_cutoff;
        st->h_is_sparse[r] = sp;
        st->h_off[r] = st->pos_off[r];
        (sp ? st->h_sparse_rows : st->h_dense_rows).push_back((uint32_t)r);
    }
    // off[r+1]-off[r] must be the list length for sparse rows: dense rows store nothing,
    // so consecutive offsets already delimit each sparse row's list.
    st->h_off[n] = c->tot_scalar;
    st->n_sparse = st->h_sparse_rows.size();
    st->n_dense = st->h_dense_rows.size();
    // every row sparse: the lists are a complete flat form of the matrix, which the row-group stream kernel can answer from
    st->stream_ok = st->h_dense_rows.empty() && stream_groups(c->n_scalar, n, &st->h_group_start);
    st->list_rows_built = n;
}

// ... and its copy on one replica.
int sync_lists(STORM_contiguous_t* c, ContigState* st, ContigDev* d) {
    build_list_meta(c, st);
    if (d->list_rows_synced == c->n_data && d->d_pos_off) return STORM_B200_OK;
    DeviceGuard guard(d->ctx.device);
    const uint64_t n = c->n_data;
    if (d->d_meta_cap < n + 1) {
        for (void* p : {(void*)d->d_pos_off, (void*)d->d_is_sparse, (void*)d->d_sparse_rows, (void*)d->d_dense_rows, (void*)d->d_group_start})
            if (p) cudaFree(p);
        d->d_pos_off = nullptr; d->d_is_sparse = nullptr; d->d_sparse_rows = nullptr; d->d_dense_rows = nullptr; d->d_group_start = nullptr;
        d->d_meta_cap = 0;                           // (a failed allocation below leaves a consistent, empty state)
        d->list_rows_synced = 0;
        const uint64_t cap = (n + 1) * 2;
        STORM_CUDA_TRY(cudaMalloc(&d->d_group_start, (cap + 1) * sizeof(uint32_t)));
        STORM_CUDA_TRY(cudaMalloc(&d->d_pos_off, cap * sizeof(uint64_t)));
        STORM_CUDA_TRY(cudaMalloc(&d->d_is_sparse, cap * sizeof(uint32_t)));
        STORM_CUDA_TRY(cudaMalloc(&d->d_sparse_rows, cap * sizeof(uint32_t)));
        STORM_CUDA_TRY(cudaMalloc(&d->d_dense_rows, cap * sizeof(uint32_t)));
        d->d_meta_cap = cap;
    }
    int rc = grow_device(&d->d_pos, &d->d_pos_cap, std::max<uint64_t>(c->tot_scalar, 1));
    if (rc) return rc;
    { int crc = copy_to_device_now(d->d_pos_off, st->h_off.data(), (n + 1) * sizeof(uint64_t)); if (crc) return crc; }
    { int crc = copy_to_device_now(d->d_is_sparse, st->h_is_sparse.data(), n * sizeof(uint32_t)); if (crc) return crc; }
    if (!st->h_sparse_rows.empty())
        { int crc = copy_to_device_now(d->d_sparse_rows, st->h_sparse_rows.data(), st->h_sparse_rows.size() * sizeof(uint32_t)); if (crc) return crc; }
    if (!st->h_dense_rows.empty())
        { int crc = copy_to_device_now(d->d_dense_rows, st->h_dense_rows.data(), st->h_dense_rows.size() * sizeof(uint32_t)); if (crc) return crc; }
    if (c->tot_scalar)
        { int crc = copy_to_device_now(d->d_pos, c->scalar, c->tot_scalar * sizeof(uint32_t)); if (crc) return crc; }
    if (st->stream_ok)
        { int crc = copy_to_device_now(d->d_group_start, st->h_group_start.data(), st->h_group_start.size() * sizeof(uint32_t)); if (crc) return crc; }
    d->list_rows_synced = n;
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
std::atomic<int> g_list_route{0};   // STORM_b200_set_contig_list_route: 0 cost model, 1 tile kernel, 2 tile + probe kernels, 3 stream kernel

// Shared body of all contiguous queries.  Replica g of G answers shard (shard * G + g) of (n_shards * G); the host adds the G totals.
uint64_t contig_query(STORM_contiguous_t* c, QueryMode mode, uint32_t shard, uint32_t n_shards, int kernel) {
    if (c == nullptr) return (uint64_t)-1;                               // storm.c:1150,1176
    ContigState* st = state_of(c);
    if (mode == QUERY_LIST) {
        if (c->scalar == nullptr) return (uint64_t)-2;                   // storm.c:1245
        if (c->n_scalar == nullptr) return (uint64_t)-3;                 // storm.c:1246
    }
    if (c->n_data < 2) return 0;
    DeviceGuard home(st->home_device);
    if (ensure_devs(st)) return (uint64_t)-1;
    const int G = (int)st->devs.size();
    const auto t0 = std::chrono::steady_clock::now();
    ContigDev* prim = st->devs[0];
    { DeviceGuard guard(prim->ctx.device); cudaEventRecord(prim->ev[0], prim->ctx.stream); }
    std::vector<DevCtx*> ctxs(G);
    std::vector<uint64_t*> arenas(G);
    for (int g = 0; g < G; ++g) {
        ContigDev* d = st->devs[g];
        if (ensure_device_rows(c, st, d, c->n_data)) return (uint64_t)-1;
        if (order_after_uploads(d)) return (uint64_t)-1;                  // rows pushed while STORM_contig_add was running
        if (d->ctx.zero_total()) return (uint64_t)-1;
        ctxs[g] = &d->ctx; arenas[g] = d->d_rows;
    }
    { DeviceGuard guard(prim->ctx.device); cudaEventRecord(prim->ev[1], prim->ctx.stream); }

    int rc = STORM_B200_OK;
    uint64_t fused = RESIDENT_TOTAL_NONE;                                 // set by banded_triangle if it has read the totals back itself
    bool hybrid = false, stream = false;
    if (mode == QUERY_LIST) {
        build_list_meta(c, st);
        // storm.c:1253-1258 switches per pair between the probe and the bitmap kernel; both give |i AND j|, so the
        // switch is a cost decision here (fitted: tile kernel 6e13 wp/s; probe kernel 0.08 ns + 1.3 ps per value
        // and pair, profiles/r01_configs_c1_c4_full.jsonl; stream kernel: stream_seconds): all rows through the
        // tile kernel, or dense x dense pairs through it and the rest probed, or -- when every row is a list --
        // the row-group stream kernel over the lists.  A pure function of the container: every shard decides alike.
        const double N = (double)c->n_data, W = (double)c->n_bitmaps_vector, pairs = 0.5 * N * (N - 1.0);
        const double avg = (double)c->tot_scalar / std::max(1.0, (double)st->n_sparse);
        const double nd = (double)st->n_dense;
        const double tile_all_s = pairs * W / 6e13 + 3e-5;
        const double hybrid_s = 0.5 * nd * (nd - 1.0) * W / 6e13 + (nd >= 2 ? 3e-5 + nd * W * 8 / 2e12 : 0.0) +
                                (pairs - 0.5 * nd * (nd - 1.0)) * (0.08e-9 + 1.3e-12 * avg) + 1e-5;
        const double stream_s = st->stream_ok ? stream_seconds(pairs, avg) : 1e30;
        const int forced = g_list_route.load();
        if (st->n_sparse > 0) {
            stream = stream_s < tile_all_s && stream_s < hybrid_s;
            hybrid = !stream && hybrid_s < tile_all_s;
            if (forced == 1) { stream = false; hybrid = false; }
            if (forced == 2) { stream = false; hybrid = true; }
            if (forced == 3) { stream = st->stream_ok; hybrid = !stream; }
        }
        st->last_list_route = stream ? 3 : hybrid ? 2 : 1;
    }
    if (!stream && !hybrid) {
        // every pair through the tile kernel: rows not yet on the devices go up band by band (1/G of a band per
        // PCIe link, the rest from the peers) while the tiles of the bands already complete are being computed
        rc = banded_triangle(ctxs.data(), arenas.data(), G, st->stride, mirror_of(c, st), st->uploaded_rows, c->n_data,
                             c->n_bitmaps_vector, shard, n_shards, kernel, &fused);
        if (!rc) st->uploaded_rows = c->n_data;
    } else {
        if (!stream && (upload_pending(c, st, c->n_data))) return (uint64_t)-1;          // the probe kernel reads the bitmaps
        for (int g = 0; g < G && !rc; ++g) {
            ContigDev* d = st->devs[g];
            const uint32_t sh = shard * (uint32_t)G + (uint32_t)g, nsh = n_shards * (uint32_t)G;
            if ((rc = sync_lists(c, st, d))) break;
            DeviceGuard guard(d->ctx.device);
            if (stream) {
                rc = launch_sparse_stream(d->d_pos_off, d->d_pos, st->h_group_start, d->d_group_start, d->d_pos_off, d->d_pos,
                                          c->tot_scalar, 0, c->n_data, 0, c->n_data, 1, sh, nsh, d->ctx.d_total, d->ctx.stream);
                continue;
            }
            if ((rc = order_after_uploads(d))) break;
            // per-pair dispatch of storm.c:1253-1258: dense x dense pairs -> tile kernel on the
            // compacted dense rows; every pair with a sparse row -> probe kernel.
            if (st->n_dense >= 2) {
                if (d->d_gather_cap < st->n_dense * st->stride) {
                    if (d->d_gather) cudaFree(d->d_gather);
                    d->d_gather = nullptr; d->d_gather_cap = 0;
                    if (cudaMalloc(&d->d_gather, st->n_dense * st->stride * 8) != cudaSuccess) {
                        cudaGetLastError(); d->d_gather = nullptr; set_error("device allocation for the dense-row gather failed"); return (uint64_t)-1;
                    }
                    d->d_gather_cap = st->n_dense * st->stride;
                }
                rc = launch_gather_rows(d->d_gather, d->d_rows, st->stride, d->d_dense_rows, st->n_dense, d->ctx.stream);
                if (!rc) rc = pairw_triangle(d->d_gather, st->n_dense, c->n_bitmaps_vector, st->stride, sh, nsh,
                                             kernel, reinterpret_cast<uint64_t*>(d->ctx.d_total), d->ctx.stream);
            }
            if (!rc) {
                uint64_t b, e;
                shard_range(st->n_sparse, sh, nsh, &b, &e);
                rc = launch_contig_probe(d->d_rows, st->stride, c->n_data, d->d_is_sparse, d->d_sparse_rows + b, e - b,
                                         d->d_pos, d->d_pos_off, d->ctx.d_total, d->ctx.stream);
            }
        }
    }
    if (rc) return (uint64_t)-1;
    { DeviceGuard guard(prim->ctx.device); cudaEventRecord(prim->ev[2], prim->ctx.stream); }
    const uint64_t total = fused != RESIDENT_TOTAL_NONE ? fused : collect_totals(ctxs.data(), G, "query");
    if (total == (uint64_t)-1) return total;
    if (fused != RESIDENT_TOTAL_NONE) { DeviceGuard guard(prim->ctx.device); cudaEventSynchronize(prim->ev[2]); }
    float up = 0, kn = 0;
    cudaEventElapsedTime(&up, prim->ev[0], prim->ev[1]);
    cudaEventElapsedTime(&kn, prim->ev[1], prim->ev[2]);
    st->timing[0] = up * 1e-3;
    st->timing[1] = kn * 1e-3;
    st->timing[2] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return total;
}

// The host mirror (public field `data`): page-locked when a device is there, so that uploads are direct DMA at
// PCIe rate and can run while STORM_contig_add is still being called; plain memory on a box without one.
uint64_t* alloc_mirror(uint64_t bytes, bool* pinned) {
    void* p = nullptr;
    *pinned = false;
    if (have_device() && cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) { *pinned = true; return (uint64_t*)p; }
    cudaGetLastError();
    if (posix_memalign(&p, 4096, bytes)) return nullptr;
    return (uint64_t*)p;
}
void free_mirror(uint64_t* p, bool pinned) {
    if (!p) return;
    if (pinned) cudaFreeHost(p); else free(p);
}

int grow_host_rows(STORM_contiguous_t* c, uint64_t need) {
    if (need <= c->m_data) return 0;
    ContigState* st = state_of(c);
    uint64_t cap = std::max<uint64_t>({need, (uint64_t)512, c->m_data * 2});     // geometric, not +512 (storm.c:1078-1100)
    const uint64_t W = c->n_bitmaps_vector;
    bool pinned = false;
    uint64_t* fresh = alloc_mirror(std::max<uint64_t>(cap * W * 8, 64), &pinned);
    if (!fresh) return -1;
    if (c->data) memcpy(fresh, c->data, c->n_data * W * 8);
    memset(fresh + c->n_data * W, 0, (cap - c->n_data) * W * 8);
    uint32_t* ns = (uint32_t*)realloc(c->n_scalar, cap * sizeof(uint32_t));
    STORM_contiguous_bitmap_t* bm = (STORM_contiguous_bitmap_t*)realloc(c->bitmaps, cap * sizeof(STORM_contiguous_bitmap_t));
    if (!ns || !bm) { free_mirror(fresh, pinned); if (ns) c->n_scalar = ns; if (bm) c->bitmaps = bm; return -1; }
    wait_uploads(st);                                                     // copies in flight still read the old mirror
    free_mirror(c->data, st->data_pinned);
    c->data = fresh; st->data_pinned = pinned; c->n_scalar = ns; c->bitmaps = bm; c->m_data = cap;
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

// ---- scratch arenas of the raw-buffer wrappers: one per device of the current device set ---------------------
struct WrapDev {
    DevCtx ctx;
    uint64_t* d_rows = nullptr; uint64_t cap_words = 0;
};
// One set of scratch arenas per device set that has been used (a caller that alternates between devices, or between
// one device and all of them, keeps its arenas; at most four sets are kept, the least recently used one goes).
struct WrapSet {
    std::vector<int> ids;
    std::vector<WrapDev*> devs;
    uint64_t last_use = 0;
};
struct WrapState {
    std::mutex mu;
    std::vector<WrapSet> sets;
    uint64_t clock = 0;
    std::vector<WrapDev*> devs;                      // = the arenas of the set in use by the current call
};
WrapState g_wrap;

void wrap_free_set(WrapSet* ws) {
    for (WrapDev* d : ws->devs) {
        if (d->ctx.device >= 0) { DeviceGuard guard(d->ctx.device); if (d->d_rows) cudaFree(d->d_rows); }
        d->ctx.destroy();
        delete d;
    }
    ws->devs.clear();
}

// Arenas of `words` words on the first n_use devices of the set (all of them for n_use <= 0).
int wrap_prepare(uint64_t words, int n_use) {
    std::vector<int> ids;
    int rc = query_devices(&ids);
    if (rc) return rc;
    WrapSet* use = nullptr;
    for (WrapSet& ws : g_wrap.sets) if (ws.ids == ids) use = &ws;
    if (!use) {
        if (g_wrap.sets.size() >= 4) {                                    // drop the least recently used set
            size_t victim = 0;
            for (size_t k = 1; k < g_wrap.sets.size(); ++k) if (g_wrap.sets[k].last_use < g_wrap.sets[victim].last_use) victim = k;
            wrap_free_set(&g_wrap.sets[victim]);
            g_wrap.sets.erase(g_wrap.sets.begin() + (long)victim);
        }
        g_wrap.sets.emplace_back();
        use = &g_wrap.sets.back();
        use->ids = ids;
        enable_peers(ids);
        for (int id : ids) {
            WrapDev* d = new (std::nothrow) WrapDev();
            if (!d) { set_error("out of host memory"); return STORM_B200_ENOMEM; }
            use->devs.push_back(d);
            if ((rc = d->ctx.init(id))) { wrap_free_set(use); g_wrap.sets.pop_back(); return rc; }
        }
    }
    use->last_use = ++g_wrap.clock;
    g_wrap.devs = use->devs;
    const int n = n_use <= 0 ? (int)g_wrap.devs.size() : std::min<int>(n_use, (int)g_wrap.devs.size());
    for (int g = 0; g < n; ++g) {
        WrapDev* d = g_wrap.devs[g];
        if (words <= d->cap_words) continue;
        DeviceGuard guard(d->ctx.device);
        if (d->d_rows) cudaFree(d->d_rows);
        d->d_rows = nullptr; d->cap_words = 0;
        if (cudaMalloc(&d->d_rows, words * 8) != cudaSuccess) {
            cudaGetLastError(); d->d_rows = nullptr; set_error("scratch arena of %llu bytes does not fit on device %d", (unsigned long long)(words * 8), d->ctx.device);
            return STORM_B200_ENOMEM;
        }
        d->cap_words = words;
    }
    return STORM_B200_OK;
}

inline uint64_t padded_stride(uint64_t n_words) { return (n_words + ROW_ALIGN_WORDS - 1) / ROW_ALIGN_WORDS * ROW_ALIGN_WORDS; }

// Upload a whole host matrix into the primary scratch arena at word offset `at`, ordered before later work on its stream.
int wrap_upload_primary(const uint64_t* vals, uint64_t n_vectors, uint64_t n_ints, uint64_t stride, uint64_t at) {
    WrapDev* d = g_wrap.devs[0];
    int rc = upload_rows(&d->ctx, d->d_rows + at, stride, HostRows{vals, n_ints, host_pointer_is_pinned(vals)}, (uint32_t)n_ints, 0, n_vectors);
    if (rc) return rc;
    DeviceGuard guard(d->ctx.device);
    STORM_CUDA_TRY(cudaEventRecord(d->ctx.mark, d->ctx.copy_stream));
    STORM_CUDA_TRY(cudaStreamWaitEvent(d->ctx.stream, d->ctx.mark, 0));
    return STORM_B200_OK;
}

// STORM_wrapper_diag[_blocked] (storm.c:132-150, 222-279) and its sharded / set-operation forms.  The host matrix
// goes up band by band over every device of the set (devices.h: banded_triangle) and the tiles that only need the
// rows already there start at once: the query costs about one band of upload plus the kernel time.
uint64_t wrapper_diag_impl(uint64_t n_vectors, const uint64_t* vals, uint64_t n_ints,
                           uint32_t shard = 0, uint32_t n_shards = 1, int kernel = STORM_B200_KERNEL_AUTO,
                           int op = STORM_B200_OP_INTERSECT) {
    if (vals == nullptr || n_ints == 0) { set_error("STORM_wrapper_diag: NULL buffer or zero width"); return (uint64_t)-1; }
    if (n_shards == 0 || shard >= n_shards) { set_error("shard %u of %u", shard, n_shards); return (uint64_t)-1; }
    if (op < 0) return (uint64_t)-1;                                     // unrecognised per-pair function (error already set)
    if (n_vectors < 2) return 0;
    std::lock_guard<std::mutex> lock(g_wrap.mu);
    const uint64_t stride = padded_stride(n_ints);
    const bool all = op == STORM_B200_OP_INTERSECT;
    if (wrap_prepare(n_vectors * stride, all ? 0 : 1)) return (uint64_t)-1;
    const int G = all ? (int)g_wrap.devs.size() : 1;
    std::vector<DevCtx*> ctxs(G);
    std::vector<uint64_t*> arenas(G);
    for (int g = 0; g < G; ++g) {
        WrapDev* d = g_wrap.devs[g];
        if (d->ctx.zero_total()) return (uint64_t)-1;
        ctxs[g] = &d->ctx; arenas[g] = d->d_rows;
    }
    if (all) {
        if (banded_triangle(ctxs.data(), arenas.data(), G, stride, HostRows{vals, n_ints, host_pointer_is_pinned(vals)}, 0, n_vectors,
                            (uint32_t)n_ints, shard, n_shards, kernel)) return (uint64_t)-1;
    } else {
        if (n_shards != 1) { set_error("set operations other than intersect are not sharded"); return (uint64_t)-1; }
        if (wrap_upload_primary(vals, n_vectors, n_ints, stride, 0)) return (uint64_t)-1;
        WrapDev* d = g_wrap.devs[0];
        DeviceGuard guard(d->ctx.device);
        if (pairw_rect_op(d->d_rows, n_vectors, stride, 0, d->d_rows, n_vectors, stride, 0, (uint32_t)n_ints, 1, op, kernel, true,
                          nullptr, 0, reinterpret_cast<uint64_t*>(d->ctx.d_total), d->ctx.stream)) return (uint64_t)-1;
    }
    return collect_totals(ctxs.data(), G, "STORM_wrapper_diag");
}

}  // namespace
}  // namespace storm

// =================================================================================
// C ABI: storm.h contiguous entry points
// =================================================================================
extern "C" {
uint64_t STORM_b200_host_intersect_count(const uint64_t* a, const uint64_t* b, const size_t n);
uint64_t STORM_b200_host_union_count(const uint64_t* a, const uint64_t* b, const size_t n);
uint64_t STORM_b200_host_diff_count(const uint64_t* a, const uint64_t* b, const size_t n);
}
namespace storm {
// Which set operation a caller's per-pair kernel stands for.  The reference's raw-buffer loops apply whatever
// STORM_compute_func they are handed (storm.c:132-150); host code cannot run on the device, so the function is
// identified instead: this library's own three by address, anything else -- libalgebra's static kernels of the
// caller's translation unit, a hand-written one -- by what it returns on three small probe vectors (and / or / xor
// counts differ on each).  NULL means intersect.  A function that is none of the three gets the error sentinel:
// answering it with an intersect total would be silently wrong.
int op_of_compute_func(const STORM_compute_func f) {
    if (f == nullptr || f == &STORM_b200_host_intersect_count) return STORM_B200_OP_INTERSECT;
    if (f == &STORM_b200_host_union_count) return STORM_B200_OP_UNION;
    if (f == &STORM_b200_host_diff_count) return STORM_B200_OP_DIFF;
    static std::mutex mu;
    static std::vector<std::pair<STORM_compute_func, int>> seen;            // a caller passes the same pointer every time
    std::lock_guard<std::mutex> lock(mu);
    for (const auto& e : seen) if (e.first == f) { if (e.second < 0) set_error("STORM_compute_func %p is neither an intersect, a union nor a diff count", (void*)f); return e.second; }
    alignas(64) uint64_t a[3][72], b[3][72];
    const size_t len[3] = {1, 7, 72};                                      // below, across and above the SIMD widths of libalgebra's kernels
    uint64_t x = 0x9E3779B97F4A7C15ull;
    auto next = [&x]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    bool is_and = true, is_or = true, is_xor = true;
    for (int t = 0; t < 3; ++t) {
        uint64_t want_and = 0, want_or = 0, want_xor = 0;
        for (size_t k = 0; k < len[t]; ++k) {
            a[t][k] = next() | next(); b[t][k] = next() & next();          // different densities: the three counts differ
            want_and += (uint64_t)__builtin_popcountll(a[t][k] & b[t][k]);
            want_or += (uint64_t)__builtin_popcountll(a[t][k] | b[t][k]);
            want_xor += (uint64_t)__builtin_popcountll(a[t][k] ^ b[t][k]);
        }
        const uint64_t got = f(a[t], b[t], len[t]);
        is_and = is_and && got == want_and; is_or = is_or && got == want_or; is_xor = is_xor && got == want_xor;
    }
    const int op = is_and ? STORM_B200_OP_INTERSECT : is_or ? STORM_B200_OP_UNION : is_xor ? STORM_B200_OP_DIFF : -1;
    if (seen.size() < 64) seen.emplace_back(f, op);
    if (op < 0) set_error("STORM_compute_func %p is neither an intersect, a union nor a diff count", (void*)f);
    return op;
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
    if (have_device()) cudaGetDevice(&st->home_device);                   // remember the creating thread's device, if any
    return c;
}

void STORM_contig_free(STORM_contiguous_t* c) {                             // storm.c:1020-1029
    if (c == nullptr) return;
    ContigState* st = state_of(c);
    bool pinned = false;
    if (st) {
        pinned = st->data_pinned;
        for (ContigDev* d : st->devs) {
            if (d->ctx.device >= 0) {
                DeviceGuard guard(d->ctx.device);
                if (d->ctx.stream) cudaStreamSynchronize(d->ctx.stream);
                if (d->ctx.copy_stream) cudaStreamSynchronize(d->ctx.copy_stream);
                for (void* p : {(void*)d->d_rows, (void*)d->d_pos, (void*)d->d_pos_off, (void*)d->d_is_sparse,
                                (void*)d->d_sparse_rows, (void*)d->d_dense_rows, (void*)d->d_group_start, (void*)d->d_gather})
                    if (p) cudaFree(p);
                for (auto e : d->ev) if (e) cudaEventDestroy(e);
            }
            d->ctx.destroy();
            delete d;
        }
        delete st;
    }
    free_mirror(c->data, pinned);
    free(c->scalar); free(c->n_scalar); free(c->bitmaps);
    free(c);
}

int STORM_contig_add(STORM_contiguous_t* c, const uint32_t* values, const uint32_t n_values) {   // storm.c:1031-1137
    if (c == nullptr) return -1;
    if (values == nullptr) return -2;
    if (n_values == 0) return 0;                                          // no row appended (D7)
    ContigState* st = state_of(c);
    if (grow_host_rows(c, c->n_data + 1)) return -3;
    if (c->scalar == nullptr && grow_host_scalar(c, st, 1)) return -3;    // reference allocates it on first add (:1037-1041)

    // :1103-1115, one pass.  The bound check the reference lacks (it writes out of bounds) is a never-taken branch; a
    // position out of range undoes the row (rows are append-only: it was all zero) and adds nothing.  The reference's
    // read-modify-write of memory per bit is a dependency chain through store forwarding (~5 cycles per bit of the same
    // word: a third of the bits set means ~20 links per word), so while the positions ascend -- what every caller passes
    // -- the bits of the current word are kept in a register, masked away without a branch when the word changes, and
    // simply stored every time: the last store of a word holds all its bits.  The first descending position hands the
    // rest of the row to the read-modify-write loop.
    uint64_t* row = c->data + c->n_data * c->n_bitmaps_vector;
    const uint64_t limit = c->vector_length;
    uint32_t dups = 0, last = 0, i = 0;
    bool reject = false;
    {
        uint64_t word = ~0ull, bits = 0;
        for (; i < n_values; ++i) {
            const uint32_t v = values[i];
            if (v >= limit) { reject = true; break; }
            if (v < last) break;
            dups += (uint32_t)(i != 0) & (uint32_t)(v == last);           // adjacent duplicates are skipped (:1106-1108)
            last = v;
            const uint64_t w = v >> 6;
            bits = (bits & (0 - (uint64_t)(w == word))) | (1ull << (v & 63));
            word = w;
            row[w] = bits;
        }
    }
    for (; i < n_values && !reject; ++i) {
        const uint32_t v = values[i];
        if (v >= limit) { reject = true; break; }
        dups += (uint32_t)(v == last);
        last = v;
        row[v >> 6] |= 1ull << (v & 63);
    }
    if (reject) {
        memset(row, 0, (size_t)c->n_bitmaps_vector * sizeof(uint64_t));
        set_error("position %u >= vector_length %llu", values[i], (unsigned long long)c->vector_length);
        return -3;
    }
    const uint32_t used = n_values - dups;
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
    // Rows are append-only and complete once this call returns, and the mirror is page-locked: every few megabytes of
    // finished rows are handed to the copy engines right away, so that the first query finds (nearly) all rows
    // resident instead of paying for the whole upload.  Failures here are not errors of the add: the query retries.
    if (st->data_pinned && st->dev_status >= 0 &&
        (c->n_data - st->uploaded_rows) * (uint64_t)c->n_bitmaps_vector * 8 >= BG_UPLOAD_BYTES) {
        DeviceGuard home(st->home_device);
        const bool first = st->dev_status == 0;
        if (ensure_devs(st) || upload_pending(c, st, c->n_data)) { if (st->dev_status == 0) st->dev_status = -1; }
        else if (first) warm_up_devices(st);
    }
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
    st->list_rows_built = (uint64_t)-1;
    for (ContigDev* d : st->devs) {
        d->list_rows_synced = 0;
        if (d->d_rows) {                                                  // on the copy stream: ordered after any upload still in flight, before the next one
            DeviceGuard guard(d->ctx.device);
            cudaMemsetAsync(d->d_rows, 0, d->d_cap_rows * st->stride * 8, d->ctx.copy_stream);
        }
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
    const int prev = g_list_route.load();
    if (route >= 0 && route <= 3) g_list_route.store(route);
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

// ---- per-pair kernel pointers (libalgebra.h:3035, 3094-3236) ----------------------------------
// These exported functions are what the STORM_get_*_count_func choosers of this library return.
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

// ---- raw-buffer wrappers (storm.c:132-369) ---------------------------------------
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
    if (op < 0) return (uint64_t)-1;
    if (!vals1 || !vals2 || n_ints == 0) { set_error("STORM_wrapper_square: NULL buffer or zero width"); return (uint64_t)-1; }
    if (n_vectors1 == 0 || n_vectors2 == 0) return 0;
    std::lock_guard<std::mutex> lock(g_wrap.mu);
    const uint64_t stride = padded_stride(n_ints);
    if (wrap_prepare(((uint64_t)n_vectors1 + n_vectors2) * stride, 1)) return (uint64_t)-1;
    WrapDev* d = g_wrap.devs[0];
    const uint64_t at2 = (uint64_t)n_vectors1 * stride;
    DeviceGuard guard(d->ctx.device);
    if (d->ctx.zero_total()) return (uint64_t)-1;
    if (wrap_upload_primary(vals1, n_vectors1, n_ints, stride, 0) || wrap_upload_primary(vals2, n_vectors2, n_ints, stride, at2)) return (uint64_t)-1;
    if (pairw_rect_op(d->d_rows, n_vectors1, stride, 0, d->d_rows + at2, n_vectors2, stride, 0, n_ints, 0, op,
                      STORM_B200_KERNEL_AUTO, false, nullptr, 0, reinterpret_cast<uint64_t*>(d->ctx.d_total), d->ctx.stream)) return (uint64_t)-1;
    DevCtx* ctx = &d->ctx;
    return collect_totals(&ctx, 1, "STORM_wrapper_square");
}

// The list wrappers take caller-built position arrays (storm.c:173-220, 281-369) and switch, per pair, between
// the bitmap kernel `f` and the list probe `fl` on n_alts < cutoff.  Both branches give |i AND j| when the lists
// describe the bitmaps (which is the contract: the driver builds them from the same draws, benchmark.cpp:765-790),
// so the query is answered from the bitmaps alone, through the same kernels as STORM_wrapper_diag.
uint64_t STORM_wrapper_diag_list(const uint32_t n_vectors, const uint64_t* STORM_RESTRICT vals, const uint32_t n_ints,
                                 const uint32_t* STORM_RESTRICT n_alts, const uint32_t* STORM_RESTRICT alt_positions,
                                 const uint32_t* STORM_RESTRICT alt_offsets, const STORM_compute_func f,
                                 const STORM_compute_lfunc fl, const uint32_t cutoff) {
    (void)n_alts; (void)alt_positions; (void)alt_offsets; (void)fl; (void)cutoff;
    return wrapper_diag_impl(n_vectors, vals, n_ints, 0, 1, STORM_B200_KERNEL_AUTO, op_of_compute_func(f));
}

uint64_t STORM_wrapper_diag_list_blocked(const uint32_t n_vectors, const uint64_t* STORM_RESTRICT vals, const uint32_t n_ints,
                                         const uint32_t* STORM_RESTRICT n_alts, const uint32_t* STORM_RESTRICT alt_positions,
                                         const uint32_t* STORM_RESTRICT alt_offsets, const STORM_compute_func f,
                                         const STORM_compute_lfunc fl, const uint32_t cutoff, uint32_t block_size) {
    (void)n_alts; (void)alt_positions; (void)alt_offsets; (void)fl; (void)cutoff; (void)block_size;
    return wrapper_diag_impl(n_vectors, vals, n_ints, 0, 1, STORM_B200_KERNEL_AUTO, op_of_compute_func(f));
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

// Number of device replicas the container's queries run on (0 before its first use of a device).
int STORM_b200_contig_device_count(const STORM_contiguous_t* c) {
    if (c == nullptr || c->b200 == nullptr) return 0;
    return (int)state_of(const_cast<STORM_contiguous_t*>(c))->devs.size();
}

int STORM_b200_contig_pairw_rect(STORM_contiguous_t* c, uint64_t i0, uint64_t i1, uint64_t j0, uint64_t j1, uint32_t* out) {
    if (c == nullptr || out == nullptr) { set_error("NULL argument"); return STORM_B200_EINVAL; }
    if (i0 > i1 || j0 > j1 || i1 > c->n_data || j1 > c->n_data) { set_error("rectangle outside the %llu rows", (unsigned long long)c->n_data); return STORM_B200_EINVAL; }
    if (i0 == i1 || j0 == j1) return STORM_B200_OK;
    ContigState* st = state_of(c);
    DeviceGuard home(st->home_device);
    int rc = ensure_devs(st);
    if (rc) return rc;
    if ((rc = upload_pending(c, st, c->n_data))) return rc;
    ContigDev* d = st->devs[0];
    if ((rc = order_after_uploads(d))) return rc;
    DeviceGuard guard(d->ctx.device);
    const uint64_t ni = i1 - i0, nj = j1 - j0;
    uint32_t* d_out = nullptr;
    if (cudaMalloc(&d_out, ni * nj * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); set_error("device allocation for %llu x %llu counts failed", (unsigned long long)ni, (unsigned long long)nj); return STORM_B200_ENOMEM; }
    rc = pairw_rect(d->d_rows + i0 * st->stride, ni, st->stride, i0, d->d_rows + j0 * st->stride, nj, st->stride, j0,
                    c->n_bitmaps_vector, 1, STORM_B200_KERNEL_AUTO, d_out, nj, nullptr, d->ctx.stream);
    if (!rc && (cudaMemcpyAsync(out, d_out, ni * nj * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->ctx.stream) != cudaSuccess ||
                cudaStreamSynchronize(d->ctx.stream) != cudaSuccess)) {
        set_error("rect query failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = STORM_B200_ECUDA;
    }
    cudaFree(d_out);
    return rc;
}

const uint64_t* STORM_b200_contig_device_rows(STORM_contiguous_t* c, uint64_t* row_stride_words) {
    if (c == nullptr) return nullptr;
    ContigState* st = state_of(c);
    DeviceGuard home(st->home_device);
    if (ensure_devs(st) || upload_pending(c, st, c->n_data)) return nullptr;
    ContigDev* d = st->devs[0];
    DeviceGuard guard(d->ctx.device);
    if (cudaStreamSynchronize(d->ctx.copy_stream) != cudaSuccess) return nullptr;
    if (row_stride_words) *row_stride_words = st->stride;
    return d->d_rows;
}

int STORM_b200_contig_invalidate_device(STORM_contiguous_t* c) {
    if (c == nullptr) return STORM_B200_EINVAL;
    ContigState* st = state_of(c);
    wait_uploads(st);
    st->uploaded_rows = 0;
    st->list_rows_built = (uint64_t)-1;
    for (ContigDev* d : st->devs) d->list_rows_synced = 0;
    return STORM_B200_OK;
}

int STORM_b200_contig_add_bulk(STORM_contiguous_t* c, const uint32_t* positions, const uint64_t* offsets, uint64_t n_rows) {
    if (c == nullptr || offsets == nullptr || (positions == nullptr && n_rows && offsets[n_rows] != offsets[0])) { set_error("NULL argument"); return STORM_B200_EINVAL; }
    if (n_rows == 0) return STORM_B200_OK;
    ContigState* st = state_of(c);
    DeviceGuard home(st->home_device);
    int rc = ensure_devs(st);
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
    if ((rc = upload_pending(c, st, c->n_data))) return rc;               // earlier rows first, on every replica
    ContigDev* d = st->devs[0];                                           // the bits are scattered on the primary device
    if ((rc = ensure_device_rows(c, st, d, c->n_data + n_new))) return rc;
    DeviceGuard guard(d->ctx.device);
    cudaStream_t cs = d->ctx.copy_stream;                                 // the stream the arena's other writes are ordered on

    struct DeviceTemp { void* p = nullptr; ~DeviceTemp() { if (p) cudaFree(p); } } t_pos, t_off;   // released on every return path
    STORM_CUDA_TRY(cudaMalloc(&t_pos.p, std::max<size_t>(n_pos, 1) * sizeof(uint32_t)));
    STORM_CUDA_TRY(cudaMalloc(&t_off.p, off.size() * sizeof(uint64_t)));
    uint32_t* d_pos = static_cast<uint32_t*>(t_pos.p);
    uint64_t* d_off = static_cast<uint64_t*>(t_off.p);
    STORM_CUDA_TRY(cudaMemcpyAsync(d_pos, src_pos, n_pos * sizeof(uint32_t), cudaMemcpyHostToDevice, cs));
    STORM_CUDA_TRY(cudaMemcpyAsync(d_off, off.data(), off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, cs));
    uint64_t* d_dst = d->d_rows + c->n_data * st->stride;
    STORM_CUDA_TRY(cudaMemsetAsync(d_dst, 0, n_new * st->stride * 8, cs));
    rc = launch_scatter_positions(d_dst, st->stride, d_pos, d_off, n_new, cs);
    // keep the public host mirror valid: copy the scattered rows back
    const uint64_t W = c->n_bitmaps_vector;
    if (!rc && cudaMemcpy2DAsync(c->data + c->n_data * W, W * 8, d_dst, st->stride * 8, W * 8, n_new,
                                 cudaMemcpyDeviceToHost, cs) != cudaSuccess) rc = STORM_B200_ECUDA;
    if (cudaStreamSynchronize(cs) != cudaSuccess) rc = STORM_B200_ECUDA;
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
    // the other replicas take the new rows from the mirror
    for (size_t g = 1; g < st->devs.size(); ++g) {
        ContigDev* o = st->devs[g];
        if ((rc = ensure_device_rows(c, st, o, c->n_data + n_new))) return rc;
        if ((rc = upload_rows(&o->ctx, o->d_rows, st->stride, mirror_of(c, st), c->n_bitmaps_vector, c->n_data, c->n_data + n_new))) return rc;
    }
    c->n_data += n_new;
    st->uploaded_rows = c->n_data;            // every replica holds them (or has the copy queued)
    return STORM_B200_OK;
}

// Rows handed over as bitmaps (the container's own layout: n_rows x W words at `pitch_words`, bit v of a row =
// word[v / 64] >> (v % 64) & 1; storm.c:1114).  Same row semantics as STORM_contig_add on the row's sorted positions: an
// all-zero row appends nothing (D7), rows below the cutoff also get their position list.  Bits at or above
// vector_length must be zero.
int STORM_b200_contig_add_dense(STORM_contiguous_t* c, const uint64_t* rows, uint64_t n_rows, uint64_t pitch_words) {
    if (c == nullptr || (rows == nullptr && n_rows)) { set_error("NULL argument"); return STORM_B200_EINVAL; }
    const uint64_t W = c->n_bitmaps_vector;
    if (pitch_words < W) { set_error("pitch of %llu words < %llu words per row", (unsigned long long)pitch_words, (unsigned long long)W); return STORM_B200_EINVAL; }
    if (n_rows == 0) return STORM_B200_OK;
    ContigState* st = state_of(c);
    const uint32_t tail_bits = (uint32_t)(c->vector_length & 63);
    const uint64_t tail_mask = tail_bits ? ~((1ull << tail_bits) - 1) : 0ull;     // bits of the last word that must be clear
    if (tail_mask)
        for (uint64_t r = 0; r < n_rows; ++r)
            if (rows[r * pitch_words + W - 1] & tail_mask) { set_error("row %llu has bits set at or above vector_length", (unsigned long long)r); return STORM_B200_EINVAL; }
    if (grow_host_rows(c, c->n_data + n_rows)) { set_error("host arena allocation failed"); return STORM_B200_ENOMEM; }
    if (c->scalar == nullptr && grow_host_scalar(c, st, 1)) return STORM_B200_ENOMEM;
    if (st->pos_off.size() < c->n_data + n_rows) st->pos_off.resize(c->n_data + n_rows);
    for (uint64_t r = 0; r < n_rows; ++r) {
        const uint64_t* src = rows + r * pitch_words;
        uint64_t used = 0;
        for (uint64_t k = 0; k < W; ++k) used += (uint64_t)__builtin_popcountll(src[k]);
        if (used == 0) continue;
        const uint64_t row = c->n_data;
        memcpy(c->data + row * W, src, W * 8);
        st->pos_off[row] = c->tot_scalar;
        c->bitmaps[row].scalar = nullptr;
        if (used < c->scalar_cutoff) {
            if (grow_host_scalar(c, st, c->tot_scalar + used)) return STORM_B200_ENOMEM;
            uint32_t* dst = c->scalar + c->tot_scalar;
            for (uint64_t k = 0; k < W; ++k)
                for (uint64_t w = src[k]; w; w &= w - 1) *dst++ = (uint32_t)(k * 64 + (uint64_t)__builtin_ctzll(w));
            c->bitmaps[row].scalar = c->scalar + c->tot_scalar;
            c->tot_scalar += used;
        }
        c->n_scalar[row] = (uint32_t)used;
        c->bitmaps[row].n_scalar = (uint32_t)used;
        ++c->n_data;
    }
    if (st->data_pinned && st->dev_status >= 0) {                          // hand the new rows to the copy engines right away
        DeviceGuard home(st->home_device);
        const bool first = st->dev_status == 0;
        if (ensure_devs(st) || upload_pending(c, st, c->n_data)) { if (st->dev_status == 0) st->dev_status = -1; }
        else if (first) warm_up_devices(st);
    }
    return STORM_B200_OK;
}

// Move the container's device replicas to the device set in force now (STORM_b200_set_devices / the calling thread's
// current device): the replicas are dropped and rebuilt from the host mirror at the next query or add.
int STORM_b200_contig_rehome(STORM_contiguous_t* c) {
    if (c == nullptr) return STORM_B200_EINVAL;
    ContigState* st = state_of(c);
    for (ContigDev* d : st->devs) {
        if (d->ctx.device >= 0) {
            DeviceGuard guard(d->ctx.device);
            if (d->ctx.stream) cudaStreamSynchronize(d->ctx.stream);
            if (d->ctx.copy_stream) cudaStreamSynchronize(d->ctx.copy_stream);
            for (void* p : {(void*)d->d_rows, (void*)d->d_pos, (void*)d->d_pos_off, (void*)d->d_is_sparse,
                            (void*)d->d_sparse_rows, (void*)d->d_dense_rows, (void*)d->d_group_start, (void*)d->d_gather})
                if (p) cudaFree(p);
            for (auto e : d->ev) if (e) cudaEventDestroy(e);
        }
        d->ctx.destroy();
        delete d;
    }
    st->devs.clear();
    st->dev_status = 0;
    st->uploaded_rows = 0;
    st->home_device = -1;
    if (have_device()) cudaGetDevice(&st->home_device);
    return STORM_B200_OK;
}

// Upper-triangle total of a matrix that is ALREADY resident on several devices (d_rows[g] on device device_ids[g],
// same n_rows / n_words / row stride everywhere): device g computes shard g of the tile raster, the host adds the
// totals.  The steady-state query of the multi-device mode without a container; UINT64_MAX on error.
uint64_t STORM_b200_pairw_devices(const uint64_t* const* d_rows, const int* device_ids, int n_devices, uint64_t n_rows,
                                  uint32_t n_words, uint64_t row_stride_words, int kernel) {
    if (d_rows == nullptr || device_ids == nullptr || n_devices < 1 || n_devices > 64) { set_error("bad device list"); return (uint64_t)-1; }
    if (n_rows < 2) return 0;
    static std::mutex mu;
    static std::vector<DevCtx*> cache;                                    // one context per device ordinal, kept for the process
    std::lock_guard<std::mutex> lock(mu);
    std::vector<DevCtx*> ctxs(n_devices);
    std::vector<uint64_t*> arenas(n_devices);
    std::vector<int> ids(device_ids, device_ids + n_devices);
    for (int g = 0; g < n_devices; ++g) {
        const int id = ids[g];
        if (id < 0 || id >= 64 || d_rows[g] == nullptr) { set_error("device %d / NULL rows", id); return (uint64_t)-1; }
        if ((int)cache.size() <= id) cache.resize(id + 1, nullptr);
        if (!cache[id]) {
            DevCtx* fresh = new (std::nothrow) DevCtx();
            if (!fresh || fresh->init(id)) { if (fresh) { fresh->destroy(); delete fresh; } return (uint64_t)-1; }
            cache[id] = fresh;
        }
        for (int o = 0; o < g; ++o) if (ids[o] == id) { set_error("device %d listed twice", id); return (uint64_t)-1; }
        ctxs[g] = cache[id];
        arenas[g] = const_cast<uint64_t*>(d_rows[g]);
        if (check_rows(d_rows[g], row_stride_words, n_words)) return (uint64_t)-1;
        if (ctxs[g]->zero_total()) return (uint64_t)-1;
    }
    uint64_t fused = RESIDENT_TOTAL_NONE;
    if (banded_triangle(ctxs.data(), arenas.data(), n_devices, row_stride_words, HostRows{}, n_rows, n_rows, n_words, 0, 1, kernel, &fused)) return (uint64_t)-1;
    if (fused != RESIDENT_TOTAL_NONE) return fused;
    return collect_totals(ctxs.data(), n_devices, "STORM_b200_pairw_devices");
}

int STORM_b200_contig_last_timing(STORM_contiguous_t* c, double out_seconds[3]) {
    if (c == nullptr || out_seconds == nullptr) return STORM_B200_EINVAL;
    ContigState* st = state_of(c);
    for (int i = 0; i < 3; ++i) out_seconds[i] = st->timing[i];
    return STORM_B200_OK;
}

}  // extern "C"
