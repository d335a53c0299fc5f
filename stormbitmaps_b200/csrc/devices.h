// devices.h -- the set of GPUs a storm.h query runs on, one context per device, and the banded
// upload + tile-range pipeline that turns a host matrix into per-device partial totals.
//
// The north star shards the N x N upper triangle over the GPUs of one box BEHIND storm.h: one process, one
// calling thread, G devices (the library issues the per-device driver calls from its own worker threads).  Every device holds all rows; device g owns shard g of the
// tile raster; the host adds G uint64 totals.  A host matrix reaches the devices in row bands: each device
// uploads 1/G of a band over its own PCIe link and pulls the other slices from its peers over NVLink
// (cudaMemcpyPeerAsync on the copy engines), and because the raster is monotone in the largest row a tile
// reads (common.cuh) a device starts on the tiles of a band as soon as that band is complete on it.
#pragma once

#include <functional>
#include <vector>

#include "common.cuh"

namespace storm {

constexpr int MAX_BANDS = 16;                        // row bands of one banded query
constexpr int STAGE_SLOTS = 4;                       // pinned staging buffers for pageable host sources
constexpr uint64_t MIN_BAND_BYTES = 32ull << 20;

struct DeviceGuard {
    int prev = -1; bool active = false;
    explicit DeviceGuard(int dev) {
        if (dev < 0) return;
        // A thread that has never called cudaSetDevice has no context current: runtime calls bind one lazily, driver
        // calls (cuTensorMapEncodeTiled) do not -- so the first guard of a thread binds even when the device "is" current.
        static thread_local bool bound = false;
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        if (prev != dev || !bound) { cudaSetDevice(dev); bound = true; active = prev != dev; }
    }
    ~DeviceGuard() { if (active) cudaSetDevice(prev); }
};

// What a query needs from one device.  Created once per container (or once per wrapper scratch) and device.
struct DevCtx {
    int device = -1;
    cudaStream_t stream = nullptr;                   // kernels
    cudaStream_t copy_stream = nullptr;              // uploads (host -> this device) run ahead of the kernels here
    cudaStream_t pull_stream = nullptr;              // peer pulls of a band (NVLink), so that they overlap the next band's upload (PCIe)
    unsigned long long* d_total = nullptr;
    unsigned long long* h_total = nullptr;           // pinned
    cudaEvent_t slice_ready[MAX_BANDS] = {};         // copy_stream: this device's own slice of band k is uploaded
    cudaEvent_t band_ready[MAX_BANDS] = {};          // pull_stream: every row of band k is on this device
    cudaEvent_t stage_done[STAGE_SLOTS] = {};        // copy_stream: the H2D out of staging slot s has finished
    cudaEvent_t mark = nullptr;                      // scratch event (stream <-> copy_stream ordering)
    bool total_zero = false;                         // d_total was zeroed on `stream` after the last read-back
    int init(int dev);
    int zero_total();                                // make d_total zero on `stream` (a no-op right after collect_totals)
    void destroy();
};

// The devices new containers and raw-buffer wrapper calls use: STORM_B200_DEVICES / STORM_b200_set_devices, or --
// by default -- the calling thread's current device alone.  Every id is checked to be an sm_100 device.
int query_devices(std::vector<int>* ids);
// cudaDeviceEnablePeerAccess between every pair of distinct devices of the list (once per pair and process).
void enable_peers(const std::vector<int>& ids);
// True if the driver can DMA from this host pointer directly (cudaHostAlloc / cudaHostRegister memory).
bool host_pointer_is_pinned(const void* p);

// fn(g) for g = 0 .. n-1, each on its own host thread (the caller takes 0; workers kept for the process): the driver
// calls a query makes per device are independent and ~12 us each, which at 8 devices is a third of a 0.26 ms query when
// one thread issues them all.  fn must make its device current itself (DeviceGuard).  Returns the first failure; its
// message becomes the caller's last error.  STORM_b200_set_device_threads(0) runs the loop on the caller instead.
int for_each_device(int n, const std::function<int(int)>& fn);

struct HostRows {
    const uint64_t* base = nullptr;                  // row r at base + r * pitch_words
    uint64_t pitch_words = 0;
    bool pinned = false;
};

// Rows [r0, r1) of a host matrix -> arena + r * stride on device `dc` (copy_stream), n_words words per row.
// Pinned sources go out as one (2-D) async copy; pageable ones through the pinned staging ring in 16 MB pieces
// that a few host threads fill, so the copy runs near PCIe rate beside the kernels instead of blocking in the driver.
int upload_rows(DevCtx* dc, uint64_t* arena, uint64_t stride, const HostRows& src, uint32_t n_words, uint64_t r0, uint64_t r1);

// Upper-triangle total of a host matrix over G devices.  Rows [0, resident) are already on (or on their way to, on
// copy_stream) every device; the rest is uploaded band by band as described above, and device g computes shard
// (shard * G + g) of (n_shards * G) of the tile raster into its own d_total (zeroed by the caller on `stream`).
// Everything is asynchronous: the caller reads the G totals back and adds them (collect_totals).  One exception, for
// short queries on several devices: if `resident_total` is given (preset to RESIDENT_TOTAL_NONE) and every row is
// resident already, launch AND read-back go out in one dispatch to the per-device threads and the sum is stored there
// -- the caller then skips collect_totals.
constexpr uint64_t RESIDENT_TOTAL_NONE = ~0ull - 7;
int banded_triangle(DevCtx* const* devs, uint64_t* const* arenas, int G, uint64_t stride, const HostRows& src,
                    uint64_t resident, uint64_t n_rows, uint32_t n_words, uint32_t shard, uint32_t n_shards, int kernel,
                    uint64_t* resident_total = nullptr);

// d_total of every device -> host, summed.  Returns UINT64_MAX (and sets the error) if any device failed.
uint64_t collect_totals(DevCtx* const* devs, int G, const char* what);

}  // namespace storm
