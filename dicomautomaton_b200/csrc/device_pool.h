// dcma_b200 -- process-lifetime cache of device allocations of the engines.
//
// A one-shot registration allocates and frees ~5.4 GB (512 x 512 x 256, fp64 fields).  cudaMalloc / cudaFree of blocks
// that size cost 10-25 ms per call on a quiet box and several hundred on a busy one (measured: 400 ms of a 1.19 s call) --
// time that the reference, which allocates with malloc, does not pay.  Engines therefore return their blocks to this
// pool instead of the driver, and the next engine of the same shape takes them back: same bytes, no driver call.
// Only exact-size matches are reused (an engine's block sizes are a function of its geometry), blocks are plain
// cudaMalloc memory (CUDA IPC and peer mappings work on them as before), and the pool holds at most
// DCMA_B200_DEVICE_POOL_MB per device (default 12288; 0 disables caching) -- larger engines go straight back to the
// driver.  dcma_b200_device_pool_trim() releases everything that is cached.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace dcma {

class DevicePool {
  public:
    static DevicePool &instance() {
        static DevicePool p;
        return p;
    }
    // the current device must be `dev`
    cudaError_t acquire(int dev, size_t bytes, void **out) {
        if (dev < 0 || dev >= kMaxDev) return cudaMalloc(out, bytes);  // outside the table: plain driver allocation
        {
            std::lock_guard<std::mutex> lk(mu_);
            auto &m = free_[dev];
            auto it = m.find(bytes);
            if (it != m.end()) {
                *out = it->second;
                m.erase(it);
                cached_[dev] -= bytes;
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(out, bytes);
        if (e == cudaErrorMemoryAllocation) {  // make room: everything cached on this device goes back to the driver
            cudaGetLastError();
            trim(dev);
            e = cudaMalloc(out, bytes);
        }
        return e;
    }
    void release(int dev, void *p, size_t bytes) {
        if (!p) return;
        if (dev >= 0 && dev < kMaxDev) {
            std::lock_guard<std::mutex> lk(mu_);
            if (limit_ > 0 && cached_[dev] + bytes <= limit_) {
                free_[dev].emplace(bytes, p);
                cached_[dev] += bytes;
                return;
            }
        }
        cudaFree(p);
    }
    // dev < 0: all devices.  Leaves the current device unchanged.
    void trim(int dev = -1) {
        std::multimap<size_t, void *> victims[kMaxDev];
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (int d = 0; d < kMaxDev; ++d)
                if (dev < 0 || dev == d) {
                    victims[d].swap(free_[d]);
                    cached_[d] = 0;
                }
        }
        int cur = 0;
        cudaGetDevice(&cur);
        for (int d = 0; d < kMaxDev; ++d) {
            if (victims[d].empty()) continue;
            cudaSetDevice(d);
            for (auto &kv : victims[d]) cudaFree(kv.second);
        }
        cudaSetDevice(cur);
    }
    size_t cached_bytes() {
        std::lock_guard<std::mutex> lk(mu_);
        size_t s = 0;
        for (size_t c : cached_) s += c;
        return s;
    }

  private:
    static constexpr int kMaxDev = 64;
    DevicePool() {
        const char *e = std::getenv("DCMA_B200_DEVICE_POOL_MB");
        const long long mb = (e && *e) ? std::atoll(e) : 12288;
        limit_ = mb > 0 ? static_cast<size_t>(mb) << 20 : 0;
    }
    ~DevicePool() {}  // the driver reclaims device memory at process exit; calling into CUDA from a static destructor is unsafe
    std::mutex mu_;
    std::multimap<size_t, void *> free_[kMaxDev];
    size_t cached_[kMaxDev] = {};
    size_t limit_ = 0;
};

// Small page-locked host blocks (the engines' 72-byte control mirror): cudaMallocHost / cudaFreeHost are system calls
// that pin and unpin pages and can stall for tens of milliseconds; a handful of such blocks is kept for the life of the
// process instead.
class PinnedSmallPool {
  public:
    static PinnedSmallPool &instance() {
        static PinnedSmallPool p;
        return p;
    }
    static constexpr size_t kBlock = 256;
    cudaError_t acquire(void **out) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (!free_.empty()) {
                *out = free_.back();
                free_.pop_back();
                return cudaSuccess;
            }
        }
        return cudaMallocHost(out, kBlock);
    }
    void release(void *p) {
        if (!p) return;
        std::lock_guard<std::mutex> lk(mu_);
        free_.push_back(p);
    }

  private:
    std::mutex mu_;
    std::vector<void *> free_;
};

// cudaGetDeviceProperties fills ~1 KB from the driver on every call (milliseconds); the two fields the engines need
inline cudaError_t cached_device_info(int dev, int *major, int *minor, int *sms, char name[256]) {
    struct Info {
        bool ok = false;
        int major = 0, minor = 0, sms = 0;
        char name[256] = {0};
    };
    static Info info[64];
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    Info &i = info[dev];
    if (!i.ok) {
        cudaDeviceProp prop;
        const cudaError_t e = cudaGetDeviceProperties(&prop, dev);
        if (e != cudaSuccess) return e;
        i.major = prop.major;
        i.minor = prop.minor;
        i.sms = prop.multiProcessorCount;
        std::snprintf(i.name, sizeof i.name, "%s", prop.name);
        i.ok = true;
    }
    *major = i.major;
    *minor = i.minor;
    *sms = i.sms;
    std::snprintf(name, 256, "%s", i.name);
    return cudaSuccess;
}

}  // namespace dcma
