// devices.cu -- device set, per-device contexts, pinned staging, and the banded multi-device triangle query
// (devices.h).  Host code only; the kernels it launches are the dense tile kernels of runtime.cu.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "devices.h"
#include "runtime.h"

namespace storm {
namespace {

// ---- device set -------------------------------------------------------------------------------------------------
enum DevMode { DEV_CURRENT = 0, DEV_ALL = 1, DEV_FIRST_K = 2, DEV_LIST = 3 };
std::mutex g_dev_mu;
int g_dev_mode = DEV_CURRENT;
int g_dev_k = 1;
std::vector<int> g_dev_list;
bool g_dev_env_read = false;
bool g_peer_done[64][64] = {};

void set_from_string_locked(const char* s) {
    if (s == nullptr || *s == 0) return;
    std::string v(s);
    if (v == "all" || v == "ALL") { g_dev_mode = DEV_ALL; return; }
    if (v.find(',') == std::string::npos) {
        const int k = atoi(v.c_str());
        if (k >= 1) { g_dev_mode = DEV_FIRST_K; g_dev_k = k; }
        return;
    }
    std::vector<int> ids;                            // explicit list "0,1,2" (an id may repeat: replicas on one device)
    size_t at = 0;
    while (at <= v.size()) {
        const size_t comma = std::min(v.find(',', at), v.size());
        if (comma > at) ids.push_back(atoi(v.substr(at, comma - at).c_str()));
        at = comma + 1;
    }
    if (!ids.empty()) { g_dev_mode = DEV_LIST; g_dev_list = ids; }
}

int check_sm100(int dev) {
    static int ok[64] = {};                          // 0 unknown, 1 sm_100, 2 not
    if (dev < 0 || dev >= 64) { set_error("device id %d out of range", dev); return STORM_B200_ENODEV; }
    if (ok[dev] == 0) {
        int major = 0;
        STORM_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        ok[dev] = major == 10 ? 1 : 2;
    }
    if (ok[dev] != 1) { set_error("device %d is not an sm_100 device; libstorm_b200 is built for sm_100a only", dev); return STORM_B200_ENODEV; }
    return STORM_B200_OK;
}

// ---- pinned staging ring for pageable host sources -----------------------------------------------------------
constexpr size_t STAGE_BYTES = 16u << 20;
struct Staging {
    std::mutex mu;
    uint8_t* buf[STAGE_SLOTS] = {};
    DevCtx* last[STAGE_SLOTS] = {};                  // whose stage_done[s] says slot s is free again
    int next = 0;
};
Staging g_stage;

// rows x width bytes from a pitched source into a compact destination, split over a few host threads
void copy_rows_parallel(uint8_t* dst, const uint8_t* src, size_t src_pitch, size_t width, uint64_t rows) {
    const size_t bytes = width * rows;
    const unsigned n_thr = bytes >= (2u << 20) ? 4u : 1u;
    auto part = [=](uint64_t a, uint64_t b) {
        if (src_pitch == width) memcpy(dst + a * width, src + a * width, (b - a) * width);
        else for (uint64_t r = a; r < b; ++r) memcpy(dst + r * width, src + r * src_pitch, width);
    };
    if (n_thr == 1) { part(0, rows); return; }
    std::thread th[3];
    for (unsigned t = 1; t < n_thr; ++t) th[t - 1] = std::thread(part, rows * t / n_thr, rows * (t + 1) / n_thr);
    part(0, rows / n_thr);
    for (unsigned t = 1; t < n_thr; ++t) th[t - 1].join();
}

// ---- one host thread per additional device ----------------------------------------------------------------------
// A query on G devices issues, per device, a kernel launch, a read-back and two stream waits: ~12 us of driver time
// each.  From one host thread that is ~90 us at G = 8 -- a third of a 0.26 ms query.  The calls to different devices
// are independent, so they go out from G threads at once: the caller takes device 0, G - 1 workers (created at the
// first multi-device query, kept for the process) the others.  A worker that has just finished a job polls for the
// next one for a few tens of microseconds before it sleeps, so the two dispatches of one query (launch, collect) and
// back-to-back queries do not pay a futex wake each.
class DevicePool {
    struct Worker {
        std::thread th;
        std::mutex mu;
        std::condition_variable cv, cv_done;
        std::atomic<uint32_t> posted{0}, done{0};
        const std::function<int(int)>* fn = nullptr;
        int arg = 0, rc = 0;
        char err[512] = {};
        bool quit = false;
    };
    std::vector<std::unique_ptr<Worker>> workers_;
    std::mutex run_mu_;                              // one dispatch at a time (two host threads may query at once)

    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    static void loop(Worker* w) {
        uint32_t seen = 0;
        for (;;) {
            const auto t0 = std::chrono::steady_clock::now();
            bool got = false;
            for (int it = 0;; ++it) {
                if (w->posted.load(std::memory_order_acquire) != seen) { got = true; break; }
                if ((it & 63) == 63 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(60)) break;
                cpu_relax();
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(w->mu);
                w->cv.wait(lk, [&] { return w->quit || w->posted.load(std::memory_order_acquire) != seen; });
                if (w->quit) return;
            }
            seen = w->posted.load(std::memory_order_acquire);
            w->err[0] = 0;
            w->rc = (*w->fn)(w->arg);
            if (w->rc) { strncpy(w->err, get_error(), sizeof(w->err) - 1); w->err[sizeof(w->err) - 1] = 0; }
            { std::lock_guard<std::mutex> lk(w->mu); w->done.store(seen, std::memory_order_release); }
            w->cv_done.notify_one();
        }
    }

public:
    // fn(0) on the calling thread, fn(1) .. fn(n - 1) on the workers, all at once; returns the first failure (its
    // message becomes the caller's last error).
    int run(int n, const std::function<int(int)>& fn) {
        if (n <= 1) return n == 1 ? fn(0) : STORM_B200_OK;
        std::lock_guard<std::mutex> lock(run_mu_);
        try {
            while ((int)workers_.size() < n - 1) {
                std::unique_ptr<Worker> w(new Worker());
                Worker* raw = w.get();
                raw->th = std::thread(loop, raw);
                workers_.push_back(std::move(w));
            }
        } catch (...) {                              // no thread to be had: the caller issues everything itself
            int rc = STORM_B200_OK;
            for (int g = 0; g < n; ++g) { const int r = fn(g); if (r && !rc) rc = r; }
            return rc;
        }
        for (int k = 1; k < n; ++k) {
            Worker* w = workers_[k - 1].get();
            w->fn = &fn; w->arg = k;
            { std::lock_guard<std::mutex> lk(w->mu); w->posted.fetch_add(1, std::memory_order_release); }
            w->cv.notify_one();
        }
        int rc = fn(0);
        for (int k = 1; k < n; ++k) {
            Worker* w = workers_[k - 1].get();
            const uint32_t want = w->posted.load(std::memory_order_relaxed);
            const auto t0 = std::chrono::steady_clock::now();
            bool got = false;
            for (int it = 0;; ++it) {
                if (w->done.load(std::memory_order_acquire) == want) { got = true; break; }
                if ((it & 63) == 63 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(60)) break;
                cpu_relax();
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(w->mu);
                w->cv_done.wait(lk, [&] { return w->done.load(std::memory_order_acquire) == want; });
            }
            if (w->rc && !rc) { rc = w->rc; set_error("%s", w->err); }
        }
        return rc;
    }
    ~DevicePool() {
        for (auto& w : workers_) {
            { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
            w->cv.notify_one();
            if (w->th.joinable()) w->th.join();
        }
    }
};
DevicePool g_pool;
std::atomic<int> g_device_threads{1};               // STORM_b200_set_device_threads(0): one host thread issues everything

}  // namespace

int for_each_device(int n, const std::function<int(int)>& fn) {
    if (n > 1 && g_device_threads.load(std::memory_order_relaxed)) return g_pool.run(n, fn);
    int rc = STORM_B200_OK;
    for (int g = 0; g < n; ++g) { const int r = fn(g); if (r && !rc) rc = r; }
    return rc;
}

int DevCtx::init(int dev) {
    device = dev;
    DeviceGuard guard(dev);
    STORM_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    STORM_CUDA_TRY(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    STORM_CUDA_TRY(cudaStreamCreateWithFlags(&pull_stream, cudaStreamNonBlocking));
    STORM_CUDA_TRY(cudaMalloc(&d_total, sizeof(unsigned long long)));
    STORM_CUDA_TRY(cudaMallocHost(&h_total, sizeof(unsigned long long)));
    for (cudaEvent_t& e : slice_ready) STORM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaEvent_t& e : band_ready) STORM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaEvent_t& e : stage_done) STORM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    STORM_CUDA_TRY(cudaEventCreateWithFlags(&mark, cudaEventDisableTiming));
    return STORM_B200_OK;
}

// d_total is accumulated into by one atomic per CTA and must start at zero.  collect_totals re-zeroes it right after
// reading it back -- while the host is still waiting for the other devices -- so that the next query of a resident
// matrix does not spend a memset call per device before its launches.
int DevCtx::zero_total() {
    if (!total_zero) { DeviceGuard guard(device); STORM_CUDA_TRY(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), stream)); }
    total_zero = false;                              // (the launches that follow dirty it)
    return STORM_B200_OK;
}

void DevCtx::destroy() {
    if (device < 0) return;
    DeviceGuard guard(device);
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    if (pull_stream) cudaStreamSynchronize(pull_stream);
    if (stream) cudaStreamSynchronize(stream);
    {
        std::lock_guard<std::mutex> lock(g_stage.mu);
        for (DevCtx*& l : g_stage.last) if (l == this) l = nullptr;       // (its copies are done: synchronised above)
    }
    for (cudaEvent_t& e : slice_ready) { if (e) cudaEventDestroy(e); e = nullptr; }
    for (cudaEvent_t& e : band_ready) { if (e) cudaEventDestroy(e); e = nullptr; }
    for (cudaEvent_t& e : stage_done) { if (e) cudaEventDestroy(e); e = nullptr; }
    if (mark) cudaEventDestroy(mark);
    if (d_total) cudaFree(d_total);
    if (h_total) cudaFreeHost(h_total);
    if (stream) cudaStreamDestroy(stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (pull_stream) cudaStreamDestroy(pull_stream);
    mark = nullptr; d_total = nullptr; h_total = nullptr; stream = nullptr; copy_stream = nullptr; pull_stream = nullptr;
    device = -1;
}

int query_devices(std::vector<int>* ids) {
    ids->clear();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return STORM_B200_ENODEV;
    }
    std::lock_guard<std::mutex> lock(g_dev_mu);
    if (!g_dev_env_read) { g_dev_env_read = true; set_from_string_locked(getenv("STORM_B200_DEVICES")); }
    switch (g_dev_mode) {
        case DEV_ALL: for (int d = 0; d < n; ++d) ids->push_back(d); break;
        case DEV_FIRST_K:
            if (g_dev_k > n) { set_error("STORM_B200_DEVICES asks for %d devices, %d visible", g_dev_k, n); return STORM_B200_ENODEV; }
            for (int d = 0; d < g_dev_k; ++d) ids->push_back(d);
            break;
        case DEV_LIST:
            for (int d : g_dev_list) {
                if (d < 0 || d >= n) { set_error("STORM_B200_DEVICES names device %d, %d visible", d, n); return STORM_B200_ENODEV; }
                ids->push_back(d);
            }
            break;
        default: {
            int cur = 0;
            STORM_CUDA_TRY(cudaGetDevice(&cur));
            ids->push_back(cur);
        }
    }
    for (int d : *ids) { int rc = check_sm100(d); if (rc) return rc; }
    return STORM_B200_OK;
}

void enable_peers(const std::vector<int>& ids) {
    std::lock_guard<std::mutex> lock(g_dev_mu);
    for (int a : ids)
        for (int b : ids) {
            if (a == b || a >= 64 || b >= 64 || g_peer_done[a][b]) continue;
            g_peer_done[a][b] = true;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) { cudaGetLastError(); continue; }   // copies then stage through the host
            DeviceGuard guard(a);
            const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess) cudaGetLastError();                     // already enabled by the application: fine
        }
}

bool host_pointer_is_pinned(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int upload_rows(DevCtx* dc, uint64_t* arena, uint64_t stride, const HostRows& src, uint32_t n_words, uint64_t r0, uint64_t r1) {
    if (r1 <= r0) return STORM_B200_OK;
    DeviceGuard guard(dc->device);
    const size_t width = (size_t)n_words * 8;
    uint64_t* dst = arena + r0 * stride;
    if (stride != n_words) STORM_CUDA_TRY(cudaMemsetAsync(dst, 0, (r1 - r0) * stride * 8, dc->copy_stream));   // padding words read as zero
    const uint64_t rows_per_piece = STAGE_BYTES / width;
    if (src.pinned || rows_per_piece == 0) {         // (a row longer than a staging slot: let the driver stage it)
        STORM_CUDA_TRY(cudaMemcpy2DAsync(dst, stride * 8, src.base + r0 * src.pitch_words, src.pitch_words * 8, width, r1 - r0,
                                         cudaMemcpyHostToDevice, dc->copy_stream));
        return STORM_B200_OK;
    }
    std::lock_guard<std::mutex> lock(g_stage.mu);
    for (uint64_t p0 = r0; p0 < r1; p0 += rows_per_piece) {
        const uint64_t p1 = std::min(r1, p0 + rows_per_piece);
        const int s = g_stage.next;
        g_stage.next = (s + 1) % STAGE_SLOTS;
        if (!g_stage.buf[s]) STORM_CUDA_TRY(cudaHostAlloc(&g_stage.buf[s], STAGE_BYTES, cudaHostAllocPortable));
        if (g_stage.last[s]) STORM_CUDA_TRY(cudaEventSynchronize(g_stage.last[s]->stage_done[s]));   // its previous upload has left the slot
        copy_rows_parallel(g_stage.buf[s], reinterpret_cast<const uint8_t*>(src.base + p0 * src.pitch_words), src.pitch_words * 8, width, p1 - p0);
        STORM_CUDA_TRY(cudaMemcpy2DAsync(arena + p0 * stride, stride * 8, g_stage.buf[s], width, width, p1 - p0,
                                         cudaMemcpyHostToDevice, dc->copy_stream));
        STORM_CUDA_TRY(cudaEventRecord(dc->stage_done[s], dc->copy_stream));
        g_stage.last[s] = dc;
    }
    return STORM_B200_OK;
}

static int collect_one(DevCtx* d, const char* what);

int banded_triangle(DevCtx* const* devs, uint64_t* const* arenas, int G, uint64_t stride, const HostRows& src,
                    uint64_t resident, uint64_t n_rows, uint32_t n_words, uint32_t shard, uint32_t n_shards, int kernel,
                    uint64_t* resident_total) {
    if (G < 1 || n_shards == 0 || shard >= n_shards) { set_error("shard %u of %u on %d devices", shard, n_shards, G); return STORM_B200_EINVAL; }
    if (n_rows < 2) return STORM_B200_OK;
    // one kernel for every device (the tile shape defines the raster the shards are ranges of)
    int resolved = -1;
    for (int g = 0; g < G; ++g) {
        DeviceGuard guard(devs[g]->device);
        const int k = resolve_kernel_for_rows(kernel, arenas[g], n_rows, n_words, stride);
        if (resolved < 0) resolved = k;
        else if (k != resolved) resolved = STORM_B200_KERNEL_UMMA;        // a device that failed the FP4 self-test: exact int8 form everywhere
    }
    const TileShape ts = tile_shape_for(resolved);
    std::vector<uint64_t> prefix;
    const uint64_t n_tiles = triangle_prefix(n_rows, ts, &prefix, nullptr, nullptr);
    const uint64_t group_rows = (uint64_t)TRI_GROUP * ts.tn;              // rows a raster group adds
    const uint64_t n_groups = prefix.size() - 1;
    std::vector<uint64_t> tb(G), te(G);
    for (int g = 0; g < G; ++g) shard_range(n_tiles, shard * (uint32_t)G + (uint32_t)g, n_shards * (uint32_t)G, &tb[g], &te[g]);

    auto launch = [&](int g, uint64_t t0, uint64_t t1) -> int {
        t0 = std::max(t0, tb[g]); t1 = std::min(t1, te[g]);
        if (t1 <= t0) return STORM_B200_OK;
        DeviceGuard guard(devs[g]->device);
        return pairw_triangle_range(arenas[g], n_rows, n_words, stride, t0, t1, resolved,
                                    reinterpret_cast<uint64_t*>(devs[g]->d_total), devs[g]->stream);
    };

    // tiles of the raster groups whose rows are all resident already: one launch per device, no waiting
    if (resident > n_rows) resident = n_rows;
    const uint64_t g_res = resident >= n_rows ? n_groups : resident / group_rows;
    if (g_res == n_groups && resident_total && G > 1 && g_device_threads.load(std::memory_order_relaxed)) {
        // everything is resident: launch and read-back in ONE dispatch to the per-device threads (only with the threads:
        // issued from one thread the read-back of device g would wait for its kernel before device g + 1 is launched)
        int rc = for_each_device(G, [&](int g) -> int { int r = launch(g, 0, prefix[g_res]); return r ? r : collect_one(devs[g], "query"); });
        if (rc) return rc;
        uint64_t total = 0;
        for (int g = 0; g < G; ++g) total += *devs[g]->h_total;
        *resident_total = total;
        return STORM_B200_OK;
    }
    { int rc = for_each_device(G, [&](int g) { return launch(g, 0, prefix[g_res]); }); if (rc) return rc; }
    if (g_res == n_groups) return STORM_B200_OK;

    // the rest: bands of whole groups, at least MIN_BAND_BYTES each, at most MAX_BANDS of them
    const uint64_t left_groups = n_groups - g_res;
    uint64_t groups_per_band = std::max<uint64_t>(1, (MIN_BAND_BYTES + group_rows * n_words * 8 - 1) / (group_rows * n_words * 8));
    groups_per_band = std::max(groups_per_band, (left_groups + MAX_BANDS - 1) / MAX_BANDS);
    int band = 0;
    for (uint64_t g0 = g_res; g0 < n_groups; g0 += groups_per_band, ++band) {
        const uint64_t g1 = std::min(n_groups, g0 + groups_per_band);
        const uint64_t r0 = std::max(resident, g0 * group_rows), r1 = std::min<uint64_t>(n_rows, g1 * group_rows);
        std::vector<uint64_t> cut(G + 1);                                 // device g uploads rows [cut[g], cut[g + 1]) of the band
        for (int g = 0; g <= G; ++g) cut[g] = r0 + (r1 > r0 ? (r1 - r0) * (uint64_t)g / (uint64_t)G : 0);
        for (int g = 0; g < G; ++g) {
            int rc = upload_rows(devs[g], arenas[g], stride, src, n_words, cut[g], cut[g + 1]);
            if (rc) return rc;
            if (G > 1) { DeviceGuard guard(devs[g]->device); STORM_CUDA_TRY(cudaEventRecord(devs[g]->slice_ready[band], devs[g]->copy_stream)); }
        }
        for (int g = 0; g < G; ++g) {
            // The peer pulls of a band run on their own stream: NVLink moves band k's slices while PCIe already carries
            // band k + 1's (on one stream the upload of the next band waited for the pulls of this one).  The pull
            // stream first waits for this device's own slice -- everything queued on the copy stream before it, arena
            // zeroing included, is then done -- so that nothing can overwrite a pulled slice.
            DevCtx* d = devs[g];
            DeviceGuard guard(d->device);
            cudaStream_t ps = G > 1 ? d->pull_stream : d->copy_stream;
            if (G > 1) STORM_CUDA_TRY(cudaStreamWaitEvent(ps, d->slice_ready[band], 0));
            for (int o = 1; o < G; ++o) {                                 // pull the other slices, nearest neighbour first (spreads the NVLink load)
                const int p = (g + o) % G;
                if (cut[p + 1] <= cut[p]) continue;
                STORM_CUDA_TRY(cudaStreamWaitEvent(ps, devs[p]->slice_ready[band], 0));
                uint64_t* dst = arenas[g] + cut[p] * stride;
                const uint64_t* from = arenas[p] + cut[p] * stride;
                const size_t bytes = (cut[p + 1] - cut[p]) * stride * 8;
                if (devs[p]->device == d->device)
                    STORM_CUDA_TRY(cudaMemcpyAsync(dst, from, bytes, cudaMemcpyDeviceToDevice, ps));
                else
                    STORM_CUDA_TRY(cudaMemcpyPeerAsync(dst, d->device, from, devs[p]->device, bytes, ps));
            }
            STORM_CUDA_TRY(cudaEventRecord(d->band_ready[band], ps));
            STORM_CUDA_TRY(cudaStreamWaitEvent(d->stream, d->band_ready[band], 0));
        }
        { int rc = for_each_device(G, [&](int g) { return launch(g, prefix[g0], prefix[g1]); }); if (rc) return rc; }
    }
    return STORM_B200_OK;
}

// One device's part of collect_totals: read the total back, re-zero it for the next query while the host still waits,
// and wait for both streams (the copy stream too: a peer may still be pulling slices out of this device's arena, and
// the caller may reuse it).
static int collect_one(DevCtx* d, const char* what) {
    DeviceGuard guard(d->device);
    const cudaError_t e0 = cudaMemcpyAsync(d->h_total, d->d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, d->stream);
    d->total_zero = cudaMemsetAsync(d->d_total, 0, sizeof(unsigned long long), d->stream) == cudaSuccess;
    const cudaError_t e1 = cudaStreamSynchronize(d->stream), e2 = cudaStreamSynchronize(d->copy_stream), e3 = cudaStreamSynchronize(d->pull_stream);
    if (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        set_error("%s failed on device %d: %s", what, d->device, cudaGetErrorString(e0 != cudaSuccess ? e0 : e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
        cudaGetLastError();
        return STORM_B200_ECUDA;
    }
    return STORM_B200_OK;
}

uint64_t collect_totals(DevCtx* const* devs, int G, const char* what) {
    // (each device from its own host thread, for_each_device)
    const int rc = for_each_device(G, [&](int g) -> int { return collect_one(devs[g], what); });
    if (rc) return (uint64_t)-1;
    uint64_t total = 0;
    for (int g = 0; g < G; ++g) total += *devs[g]->h_total;
    return total;
}

}  // namespace storm

using namespace storm;

extern "C" {

// n = 0: every visible device; n >= 1: devices 0 .. n-1.  Applies to containers created and raw-buffer wrapper
// calls made afterwards.  Returns the number of devices queries used before the call (1 = the current device only).
int STORM_b200_set_devices(int n) {
    std::lock_guard<std::mutex> lock(g_dev_mu);
    if (!g_dev_env_read) { g_dev_env_read = true; set_from_string_locked(getenv("STORM_B200_DEVICES")); }
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess) { cudaGetLastError(); visible = 0; }
    const int prev = g_dev_mode == DEV_ALL ? visible : g_dev_mode == DEV_FIRST_K ? g_dev_k : g_dev_mode == DEV_LIST ? (int)g_dev_list.size() : 1;
    if (n == 0) g_dev_mode = DEV_ALL;
    else if (n >= 1) { g_dev_mode = DEV_FIRST_K; g_dev_k = n; }
    return prev;
}

// An explicit list of device ordinals (an ordinal may repeat: several replicas on one device, which is how the
// multi-device logic is exercised on a one-GPU box); n = 0 returns to the default, the calling thread's current device.
int STORM_b200_set_device_list(const int* ids, int n) {
    std::lock_guard<std::mutex> lock(g_dev_mu);
    g_dev_env_read = true;
    if (n <= 0 || ids == nullptr) { g_dev_mode = DEV_CURRENT; g_dev_list.clear(); return STORM_B200_OK; }
    g_dev_mode = DEV_LIST;
    g_dev_list.assign(ids, ids + n);
    return STORM_B200_OK;
}

// 1 (default): the per-device calls of a multi-device query go out from one host thread per device; 0: from the
// calling thread alone.  Returns the previous value.
int STORM_b200_set_device_threads(int on) { return g_device_threads.exchange(on != 0); }

// Test hook (no device needed): n jobs through the per-device thread pool; job k records that it ran, job `fail_at`
// (if >= 0) fails with a message.  Returns the pool's return code, or -100 if a job did not run exactly once.
int STORM_b200_selftest_device_threads(int n, int fail_at) {
    if (n < 0 || n > 64) return STORM_B200_EINVAL;
    std::vector<int> ran(n > 0 ? n : 1, 0);
    const int rc = for_each_device(n, [&](int k) -> int {
        ++ran[k];
        if (k == fail_at) { set_error("job %d failed on purpose", k); return STORM_B200_EINVAL; }
        return STORM_B200_OK;
    });
    for (int k = 0; k < n; ++k) if (ran[k] != 1) return -100;
    return rc;
}

// The devices a query made now would use: fills ids[0 .. min(cap, count)) and returns the count (negative on error).
int STORM_b200_get_devices(int* ids, int cap) {
    std::vector<int> v;
    int rc = query_devices(&v);
    if (rc) return rc;
    for (int i = 0; i < cap && i < (int)v.size(); ++i) ids[i] = v[i];
    return (int)v.size();
}

}  // extern "C"
