// dcma_b200 -- z-slab group: one registration sharded over several engines (see slab.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include <memory>
#include <utility>
#include <vector>

#include "engine.h"

namespace dcma {

// contiguous balanced z-slabs (earlier ranks take the remainder) extended by `halo` planes on interior sides
SlabSpec slab_for_rank(int64_t slices, int world, int rank, int64_t halo);
void nccl_unique_id(unsigned char id[128]);
// The devices a one-process registration with n_devices = n is sharded over: 0..n-1, or the first n entries of the
// comma-separated list in DCMA_B200_SLAB_DEVICES (site policy, e.g. "4,5,6,7"; repeats are allowed, which is also how
// the sharded paths are exercised on a single-GPU box).
std::vector<int> slab_devices(int n);

class SlabGroup {
  public:
    // one process, one engine per entry of `devices` (repeats allowed), halos by peer copies
    SlabGroup(const dcma_b200_geom &g, const dcma_b200_demons_params &p, const std::vector<int> &devices, int64_t halo);
    // one rank of `world` processes, halos by NCCL send/recv
    SlabGroup(const dcma_b200_geom &g, const dcma_b200_demons_params &p, const unsigned char id[128], int rank, int world,
              int device, int64_t halo);
    ~SlabGroup();
    void load(const float *fixed_whole, const float *moving_whole, bool on_device);
    void run(int64_t max_iterations, double *mse_history_host, dcma_b200_run_stats *stats);
    void owned_range(int64_t *z0, int64_t *z1) const;  // slices held by THIS process
    void get_deformation(double *host_aos);             // those slices, AoS doubles
    // warm start (pyramid): D := the given field of the WHOLE volume, a device buffer every engine of the group can read
    void import_deformation_device(const double *device_aos_whole);
    int64_t device_bytes() const;
    int64_t halo() const { return halo_; }
    int64_t halo_bytes_last_run() const { return halo_bytes_; }
    double exchange_ms_last_run() const { return exchange_ms_; }    // device time inside the exchanges (all of them)
    int64_t exchanges_last_run() const { return exchanges_; }
    bool overlapped_last_run() const { return overlapped_; }
    // 0: in-process peer copies + events, 1: NCCL send/recv, 2: IPC pull + NCCL tokens, 3: IPC push + flag counters
    int transport() const { return !comm_ ? 0 : (flags_ok_ ? 3 : (ipc_ok_ ? 2 : 1)); }
    static int64_t default_halo(const dcma_b200_geom &g, const dcma_b200_demons_params &p);

  private:
    struct Local;
    void add_local(const dcma_b200_geom &g, const dcma_b200_demons_params &p, int rank, int device);
    void exchange(int buffer, int64_t planes);
    void gather_mse_and_stop_rule(long long it);
    dcma_b200_geom geom_;
    dcma_b200_demons_params params_;
    int world_ = 1;
    int64_t halo_ = 0;
    int64_t halo_bytes_ = 0;
    void *comm_ = nullptr;  // ncclComm_t
    // NCCL groups on NVLink: the neighbours' field allocations opened through CUDA IPC, so that the halo planes move
    // by copy-engine transfers (no SMs, full NVLink bandwidth) and NCCL only carries 4-byte ordering tokens
    void setup_ipc();
    void token_sync();
    bool ipc_ok_ = false;
    static constexpr int kMaxFieldAllocs = 4;  // D, U, scratch (+ U's own scratch in DCMA_B200_FIELD_MIXED)
    void *peer_lo_[kMaxFieldAllocs] = {nullptr, nullptr, nullptr, nullptr}, *peer_hi_[kMaxFieldAllocs] = {nullptr, nullptr, nullptr, nullptr};
    int *d_token_ = nullptr;
    // flag transport: counters raised by the neighbours in my flag block, waited for with stream memory operations
    void exchange_flags(int buffer, int64_t planes, cudaStream_t st);
    bool flags_ok_ = false, overlapped_ = false;
    unsigned *d_flags_ = nullptr, *peer_flags_lo_ = nullptr, *peer_flags_hi_ = nullptr;
    unsigned xid_ = 0;
    cudaStream_t comm_stream_ = nullptr;
    cudaEvent_t ev_main_ = nullptr, ev_x_[3] = {nullptr, nullptr, nullptr};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> x_events_;
    double exchange_ms_ = 0.0;
    int64_t exchanges_ = 0;
    std::vector<std::unique_ptr<Local>> locals_;
};

}  // namespace dcma
