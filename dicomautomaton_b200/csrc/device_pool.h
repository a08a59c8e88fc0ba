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
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

namespace dcma {

class DevicePool {
  public:
    static DevicePool &instance() {
        static DevicePool p;
        return p;
    }
    // the current device must be `dev`
    cudaError_t acquire(int dev, size_t bytes, void **out) {
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
        {
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

}  // namespace dcma
