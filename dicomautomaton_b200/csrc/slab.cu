// dcma_b200 -- z-slab sharding of ONE registration over several GPUs (BASELINE config 4).
//
// The reference holds the whole volume in one address space (src/Alignment_Demons.cc:171-187) and has no multi-device
// path; this file is the B200 analogue of its loop (cc:980-995) for a volume cut into contiguous z-slabs:
//   * every slab engine holds its owned slices plus `halo` planes on each interior side; the moving image is replicated
//     (the warp gathers at p + D(p) have no a-priori bound);
//   * per iteration the neighbours exchange  r_z planes of U   before smooth(U),
//                                            halo planes of U~ before the diffeomorphic composition (its gathers),
//                                            r_z planes of D   before smooth(D),
//     and all slabs gather their (sum of squared differences, valid count) so that each applies the reference's stop
//     rule to the same totals (summed in rank order: every rank takes the same decision bit for bit);
//   * arithmetic per voxel is exactly that of the whole-volume engine: a slab run reproduces its field bit for bit.
// Two transports: peer copies between engines of ONE process (any mix of devices; also the single-GPU test vehicle),
// and NCCL send/recv + all-gather between processes (one rank per GPU, the bench/torchrun layout).  NCCL is loaded
// with dlopen so that the library itself has no link-time dependency on it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"
#include "slab.h"

namespace dcma {

SlabSpec slab_for_rank(int64_t slices, int world, int rank, int64_t halo) {
    if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("bad world/rank");
    if (halo < 0) throw std::invalid_argument("negative halo");
    const int64_t base = slices / world, rem = slices % world;
    if (base < 1) throw std::invalid_argument("more slabs than slices");
    if (world > 1 && base < halo) throw std::invalid_argument("slabs thinner than the halo: use fewer devices or a smaller halo");
    const int64_t z0 = rank * base + std::min<int64_t>(rank, rem);
    const int64_t z1 = z0 + base + (rank < rem ? 1 : 0);
    const int64_t lo = std::max<int64_t>(0, z0 - halo), hi = std::min<int64_t>(slices, z1 + halo);
    SlabSpec s;
    s.zoff = lo;
    s.local = hi - lo;
    s.own0 = z0 - lo;
    s.own1 = z1 - lo;
    return s;
}

// ---------------------------------------------------------------------------------------------------------------
// NCCL through dlopen
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

NcclApi &nccl() {
    static NcclApi api;
    if (api.handle) return api;
    const char *names[] = {std::getenv("DCMA_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) throw std::runtime_error("NCCL (libnccl.so.2) could not be loaded; set DCMA_B200_NCCL_LIB to its path");
#define DCMA_NCCL_SYM(field, sym)                                                              \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, #sym));               \
    if (!api.field) throw std::runtime_error(std::string("NCCL symbol missing: ") + #sym)
    DCMA_NCCL_SYM(GetUniqueId, ncclGetUniqueId);
    DCMA_NCCL_SYM(CommInitRank, ncclCommInitRank);
    DCMA_NCCL_SYM(CommDestroy, ncclCommDestroy);
    DCMA_NCCL_SYM(Send, ncclSend);
    DCMA_NCCL_SYM(Recv, ncclRecv);
    DCMA_NCCL_SYM(GroupStart, ncclGroupStart);
    DCMA_NCCL_SYM(GroupEnd, ncclGroupEnd);
    DCMA_NCCL_SYM(AllGather, ncclAllGather);
    DCMA_NCCL_SYM(GetErrorString, ncclGetErrorString);
#undef DCMA_NCCL_SYM
    return api;
}

#define DCMA_NCCL(expr)                                                                                            \
    do {                                                                                                           \
        ncclResult_t _r = (expr);                                                                                  \
        if (_r != ncclSuccess && _r != ncclInProgress)                                                             \
            throw std::runtime_error(std::string(#expr) + " failed: " + nccl().GetErrorString(_r));               \
    } while (0)
}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// Ordering between ranks without NCCL launches: 32-bit counters in device memory, mapped into the neighbours through
// CUDA IPC.  A rank RAISES a counter in its neighbour's memory with a one-thread kernel (a system-scope store after a
// system fence; it runs in stream order, i.e. after the copies enqueued before it have landed) and WAITS for its own
// counters with cuStreamWaitValue32 -- a front-end semaphore acquire, no SM is occupied and no kernel spins.
// ---------------------------------------------------------------------------------------------------------------
namespace {
__global__ void k_raise_flags(unsigned *a, unsigned *b, unsigned value) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        __threadfence_system();
        if (a) *reinterpret_cast<volatile unsigned *>(a) = value;
        if (b) *reinterpret_cast<volatile unsigned *>(b) = value;
        __threadfence_system();
    }
}
// fallback when stream memory operations are not available: a one-thread kernel polling a counter that only a PEER
// GPU (or a copy engine) raises, never a kernel of this device
__global__ void k_wait_flag(const unsigned *f, unsigned value) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        while (static_cast<int>(*reinterpret_cast<const volatile unsigned *>(f) - value) < 0) __nanosleep(200);
        __threadfence_system();
    }
}
// Small exchanges are latency-bound (two flag kernels + six copy-engine launches cost more than the bytes): ONE kernel does
// the whole producer side -- raise `arrived` in the neighbours, wait for theirs (a counter only a PEER GPU raises), store
// my boundary planes straight into the neighbours' halo planes over NVLink (16-byte stores), and the last block to
// finish raises `pushed`.
struct PushSeg {
    const uint4 *src;
    uint4 *dst;
    long long n16;  // 16-byte units
};
struct PushArgs {
    PushSeg seg[6];
    int nseg;
    unsigned *peer_arrived[2], *peer_pushed[2];  // counters in the neighbours' flag blocks (null: no such neighbour)
    const unsigned *my_arrived[2];                // counters in my flag block
    unsigned *ticket;                             // in my flag block
    unsigned id;
};
__global__ void __launch_bounds__(256) k_push_halo(const PushArgs a) {
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            __threadfence_system();
            for (int i = 0; i < 2; ++i)
                if (a.peer_arrived[i]) *reinterpret_cast<volatile unsigned *>(a.peer_arrived[i]) = a.id;
        }
        for (int i = 0; i < 2; ++i)
            if (a.my_arrived[i])
                while (static_cast<int>(*reinterpret_cast<const volatile unsigned *>(a.my_arrived[i]) - a.id) < 0) __nanosleep(100);
        __threadfence_system();
    }
    __syncthreads();
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (int sg = 0; sg < a.nseg; ++sg) {
        const uint4 *__restrict__ src = a.seg[sg].src;
        uint4 *__restrict__ dst = a.seg[sg].dst;
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.seg[sg].n16; i += stride) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(a.ticket, 1u);
        if (t == gridDim.x - 1) {
            *a.ticket = 0;
            __threadfence_system();
            for (int i = 0; i < 2; ++i)
                if (a.peer_pushed[i]) *reinterpret_cast<volatile unsigned *>(a.peer_pushed[i]) = a.id;
            __threadfence_system();
        }
    }
}
using StreamWaitValue32Fn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWaitValue32Fn stream_wait_value32() {
    static StreamWaitValue32Fn fn = []() -> StreamWaitValue32Fn {
        const char *off = std::getenv("DCMA_B200_SLAB_NO_MEMOPS");
        if (off && *off == '1') return nullptr;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<StreamWaitValue32Fn>(p);
    }();
    return fn;
}
void wait_flag(cudaStream_t st, const unsigned *flag, unsigned value) {
    if (StreamWaitValue32Fn fn = stream_wait_value32()) {
        const CUresult r = fn(reinterpret_cast<CUstream>(st), reinterpret_cast<CUdeviceptr>(flag), value, CU_STREAM_WAIT_VALUE_GEQ);
        if (r == CUDA_SUCCESS) return;
    }
    k_wait_flag<<<1, 32, 0, st>>>(flag, value);
    DCMA_CUDA(cudaGetLastError());
}
}  // namespace

std::vector<int> slab_devices(int n) {
    std::vector<int> dev;
    if (const char *e = std::getenv("DCMA_B200_SLAB_DEVICES")) {
        const char *p = e;
        while (*p && static_cast<int>(dev.size()) < n) {
            char *end = nullptr;
            const long v = std::strtol(p, &end, 10);
            if (end == p || v < 0 || v > 1023) throw std::invalid_argument(std::string("DCMA_B200_SLAB_DEVICES='") + e + "' is not a list of device ordinals");
            dev.push_back(static_cast<int>(v));
            p = (*end == ',') ? end + 1 : end;
            if (*end != ',' && *end != '\0') throw std::invalid_argument(std::string("DCMA_B200_SLAB_DEVICES='") + e + "' is not a list of device ordinals");
        }
        if (static_cast<int>(dev.size()) < n) throw std::invalid_argument("DCMA_B200_SLAB_DEVICES lists fewer devices than n_devices");
        return dev;
    }
    for (int i = 0; i < n; ++i) dev.push_back(i);
    return dev;
}

void nccl_unique_id(unsigned char id[128]) {
    ncclUniqueId u;
    DCMA_NCCL(nccl().GetUniqueId(&u));
    static_assert(sizeof(u.internal) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(id, u.internal, 128);
}

// ---------------------------------------------------------------------------------------------------------------
// SlabGroup
// ---------------------------------------------------------------------------------------------------------------
static inline char *plane_ptr(EngineBase *e, int buffer, int c, int64_t l0);

struct SlabGroup::Local {
    std::unique_ptr<EngineBase> eng;
    int rank = 0, device = 0;
    double *d_mse2 = nullptr;    // this slab's (sum, count)
    double *d_mse_hist = nullptr;  // per iteration (sum, count) of this slab, when no per-iteration gather is needed
    double *d_gather = nullptr;  // world x (sum, count)
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
};

SlabGroup::~SlabGroup() {
    for (auto &l : locals_) {
        cudaSetDevice(l->device);
        if (l->eng) cudaStreamSynchronize(static_cast<cudaStream_t>(l->eng->stream_handle()));
        cudaFree(l->d_mse2);
        cudaFree(l->d_mse_hist);
        cudaFree(l->d_gather);
        if (l->ev_ready) cudaEventDestroy(l->ev_ready);
        if (l->ev_done) cudaEventDestroy(l->ev_done);
        l->eng.reset();
    }
    for (int i = 0; i < kMaxFieldAllocs; ++i) {
        if (peer_lo_[i]) cudaIpcCloseMemHandle(peer_lo_[i]);
        if (peer_hi_[i]) cudaIpcCloseMemHandle(peer_hi_[i]);
    }
    cudaFree(d_token_);
    if (peer_flags_lo_) cudaIpcCloseMemHandle(peer_flags_lo_);
    if (peer_flags_hi_) cudaIpcCloseMemHandle(peer_flags_hi_);
    cudaFree(d_flags_);
    for (cudaEvent_t e : {ev_main_, ev_x_[0], ev_x_[1], ev_x_[2]})
        if (e) cudaEventDestroy(e);
    for (auto &pr : x_events_) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    if (comm_stream_) cudaStreamDestroy(comm_stream_);
    if (comm_) {
        try {
            nccl().CommDestroy(static_cast<ncclComm_t>(comm_));
        } catch (...) {
        }
    }
}

int64_t SlabGroup::default_halo(const dcma_b200_geom &g, const dcma_b200_demons_params &p) {
    auto rz = [&](double sigma) -> int64_t {
        return sigma > 0.0 ? std::max<int64_t>(1, static_cast<int64_t>(3.0 * sigma / g.pxl_dz)) : 0;  // cc:574-576
    };
    const int64_t r = std::max(rz(p.deformation_field_smoothing_sigma), p.use_diffeomorphic ? rz(p.update_field_smoothing_sigma) : 0);
    // composition gathers reach ceil(|D_z| / pxl_dz) + 1 planes beyond a slab edge: 4 planes cover |D_z| < 3 voxels (the
    // per-iteration update is clamped to max_update_magnitude and the field is smoothed, so clinical fields stay well
    // inside that).  A gather that does need a plane the slab does not hold raises Ctrl::halo_overflow and the run FAILS
    // (never silent); the one-shot call then repeats the registration with a doubled halo.
    return std::max<int64_t>(r, p.use_diffeomorphic ? 4 : 1);
}

void SlabGroup::add_local(const dcma_b200_geom &g, const dcma_b200_demons_params &p, int rank, int device) {
    // the fixed-image gradient reads the z neighbours of every owned slice (cc:254-265): at least one halo plane
    if (world_ > 1 && halo_ < 1) throw std::invalid_argument("halo_planes must be >= 1 when the volume is sharded");
    auto l = std::make_unique<Local>();
    l->rank = rank;
    dcma_b200_demons_params pp = p;
    pp.device = device;
    const SlabSpec spec = slab_for_rank(g.slices, world_, rank, halo_);
    l->eng.reset(make_engine(g, pp, nullptr, spec));
    l->device = l->eng->device();
    DCMA_CUDA(cudaSetDevice(l->device));
    DCMA_CUDA(cudaMalloc(&l->d_mse2, 2 * sizeof(double)));
    DCMA_CUDA(cudaMalloc(&l->d_gather, 2 * sizeof(double) * world_));
    DCMA_CUDA(cudaEventCreateWithFlags(&l->ev_ready, cudaEventDisableTiming));
    DCMA_CUDA(cudaEventCreateWithFlags(&l->ev_done, cudaEventDisableTiming));
    locals_.push_back(std::move(l));
}

SlabGroup::SlabGroup(const dcma_b200_geom &g, const dcma_b200_demons_params &p, const std::vector<int> &devices, int64_t halo)
    : geom_(g), params_(p), world_(static_cast<int>(devices.size())), halo_(halo < 0 ? default_halo(g, p) : halo) {
    if (devices.empty()) throw std::invalid_argument("no devices");
    for (int r = 0; r < world_; ++r) add_local(g, p, r, devices[r]);
    // peer access makes the halo copies direct NVLink transfers; failure only means staged copies
    for (auto &a : locals_)
        for (auto &b : locals_)
            if (a->device != b->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, a->device, b->device);
                if (can) {
                    cudaSetDevice(a->device);
                    if (cudaDeviceEnablePeerAccess(b->device, 0) != cudaSuccess) cudaGetLastError();
                }
            }
}

SlabGroup::SlabGroup(const dcma_b200_geom &g, const dcma_b200_demons_params &p, const unsigned char id[128], int rank, int world,
                     int device, int64_t halo)
    : geom_(g), params_(p), world_(world), halo_(halo < 0 ? default_halo(g, p) : halo) {
    if (world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("bad world/rank");
    add_local(g, p, rank, device);
    DCMA_CUDA(cudaSetDevice(locals_[0]->device));
    ncclUniqueId u;
    std::memcpy(u.internal, id, 128);
    ncclComm_t c = nullptr;
    DCMA_NCCL(nccl().CommInitRank(&c, world, u, rank));
    comm_ = c;
    const char *no_ipc = std::getenv("DCMA_B200_SLAB_NO_IPC");
    if (world > 1 && !(no_ipc && *no_ipc == '1')) setup_ipc();
}

// Exchange the CUDA IPC handles of the three field allocations with every rank (all-gather over NCCL) and open the two
// z-neighbours'.  Any failure (IPC not permitted in this environment, no peer access) leaves ipc_ok_ false on ALL ranks
// and the exchange falls back to ncclSend/ncclRecv.
void SlabGroup::setup_ipc() {
    Local &l = *locals_[0];
    EngineBase *e = l.eng.get();
    DCMA_CUDA(cudaSetDevice(l.device));
    cudaStream_t st = static_cast<cudaStream_t>(e->stream_handle());
    struct Rec {
        cudaIpcMemHandle_t h[kMaxFieldAllocs + 1];  // the field allocations (3, or 4 in mixed mode) + the flag block (last)
        int ok;
        int pad[3];
    };
    DCMA_CUDA(cudaMalloc(&d_flags_, 16 * sizeof(unsigned)));
    DCMA_CUDA(cudaMemsetAsync(d_flags_, 0, 16 * sizeof(unsigned), st));
    Rec mine;
    std::memset(&mine, 0, sizeof(mine));
    mine.ok = 1;
    const int nf = e->field_allocation_count();
    if (nf > kMaxFieldAllocs) throw std::logic_error("more field allocations than the slab group can map");
    for (int i = 0; i <= kMaxFieldAllocs; ++i) {
        if (i >= nf && i < kMaxFieldAllocs) continue;
        if (cudaIpcGetMemHandle(&mine.h[i], i < nf ? e->field_allocation(i) : static_cast<void *>(d_flags_)) != cudaSuccess) {
            cudaGetLastError();
            mine.ok = 0;
        }
    }
    Rec *d_all = nullptr, *d_mine = nullptr;
    DCMA_CUDA(cudaMalloc(&d_all, sizeof(Rec) * world_));
    DCMA_CUDA(cudaMalloc(&d_mine, sizeof(Rec)));
    DCMA_CUDA(cudaMalloc(&d_token_, 4 * sizeof(int)));
    DCMA_CUDA(cudaMemsetAsync(d_token_, 0, 4 * sizeof(int), st));
    DCMA_CUDA(cudaMemcpyAsync(d_mine, &mine, sizeof(Rec), cudaMemcpyHostToDevice, st));
    DCMA_NCCL(nccl().AllGather(d_mine, d_all, sizeof(Rec), ncclChar, static_cast<ncclComm_t>(comm_), st));
    std::vector<Rec> all(world_);
    DCMA_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(Rec) * world_, cudaMemcpyDeviceToHost, st));
    DCMA_CUDA(cudaStreamSynchronize(st));
    int ok = 1;
    for (auto &r : all) ok &= r.ok;
    auto open_peer = [&](int peer, void **out, unsigned **flags) {
        for (int i = 0; i <= kMaxFieldAllocs && ok; ++i) {
            if (i >= nf && i < kMaxFieldAllocs) continue;
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[peer].h[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
            } else if (i < nf) {
                out[i] = ptr;
            } else {
                *flags = static_cast<unsigned *>(ptr);
            }
        }
    };
    if (ok && l.rank > 0) open_peer(l.rank - 1, peer_lo_, &peer_flags_lo_);
    if (ok && l.rank + 1 < world_) open_peer(l.rank + 1, peer_hi_, &peer_flags_hi_);
    // every rank must take the same path: agree on the outcome
    mine.ok = ok;
    DCMA_CUDA(cudaMemcpyAsync(d_mine, &mine, sizeof(Rec), cudaMemcpyHostToDevice, st));
    DCMA_NCCL(nccl().AllGather(d_mine, d_all, sizeof(Rec), ncclChar, static_cast<ncclComm_t>(comm_), st));
    DCMA_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(Rec) * world_, cudaMemcpyDeviceToHost, st));
    DCMA_CUDA(cudaStreamSynchronize(st));
    ok = 1;
    for (auto &r : all) ok &= r.ok;
    ipc_ok_ = (ok != 0);
    cudaFree(d_all);
    cudaFree(d_mine);
    const char *tok = std::getenv("DCMA_B200_SLAB_NCCL_TOKENS");
    flags_ok_ = ipc_ok_ && !(tok && *tok == '1');
    if (flags_ok_) {
        // HIGH priority: the block scheduler serves a kernel of another stream only after every block of the kernels
        // launched before it has been issued -- unless its stream has priority.  Without it the one-thread flag kernels of
        // an exchange sit behind the whole interior kernel they are meant to run under (measured: no overlap at all).
        int prio_lo = 0, prio_hi = 0;
        DCMA_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        DCMA_CUDA(cudaStreamCreateWithPriority(&comm_stream_, cudaStreamNonBlocking, prio_hi));
        DCMA_CUDA(cudaEventCreateWithFlags(&ev_main_, cudaEventDisableTiming));
        for (auto &ev : ev_x_) DCMA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
}

// One halo exchange of `n` planes of `buffer` with both z-neighbours on stream `st`, PUSH model:
//   raise `arrived` in the neighbours  (everything enqueued on `st` before this point is done: nobody here still reads
//                                       the halo planes the neighbours are about to overwrite)
//   wait for the neighbours' `arrived`
//   copy my boundary planes into the neighbours' halo planes (copy engines over NVLink, no SMs)
//   raise `pushed` in the neighbours, wait for theirs: my halo planes are complete
// Counter slots in MY flag block: 0/1 = arrived from the lower/upper neighbour, 2/3 = pushed from the lower/upper one.
void SlabGroup::exchange_flags(int buffer, int64_t n, cudaStream_t st) {
    Local &l = *locals_[0];
    EngineBase *e = l.eng.get();
    const SlabSpec s = e->slab();
    const int64_t pe = e->plane_elems(), es = e->field_elem_size(buffer);
    const size_t bytes = static_cast<size_t>(n) * pe * es;
    const bool lo = (l.rank > 0), hi = (l.rank + 1 < world_);
    DCMA_CUDA(cudaSetDevice(l.device));
    int idx = -1;  // which of the field allocations currently plays `buffer` (the roles rotate identically on all ranks)
    for (int i = 0; i < e->field_allocation_count(); ++i)
        if (e->field_allocation(i) == e->field_base(buffer)) idx = i;
    if (idx < 0) throw std::logic_error("field buffer is not one of the engine's allocations");
    const unsigned id = ++xid_;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (x_events_.size() < 4096) {
        DCMA_CUDA(cudaEventCreate(&t0));
        DCMA_CUDA(cudaEventCreate(&t1));
        DCMA_CUDA(cudaEventRecord(t0, st));
    }
    static const long long kernel_push_max = []() {
        const char *v = std::getenv("DCMA_B200_SLAB_KERNEL_PUSH_MAX_MB");
        return (v && *v) ? std::atoll(v) * (1LL << 20) : 48LL * (1LL << 20);
    }();
    if (static_cast<long long>(3 * bytes) <= kernel_push_max && bytes % 16 == 0 && (pe * es) % 16 == 0) {
        PushArgs a;
        std::memset(&a, 0, sizeof(a));
        a.id = id;
        a.ticket = d_flags_ + 8;
        if (lo) {
            const SlabSpec p = slab_for_rank(geom_.slices, world_, l.rank - 1, halo_);
            for (int c = 0; c < 3; ++c) {
                PushSeg &sg = a.seg[a.nseg++];
                sg.src = reinterpret_cast<const uint4 *>(plane_ptr(e, buffer, c, s.own0));
                sg.dst = reinterpret_cast<uint4 *>(static_cast<char *>(peer_lo_[idx]) + (c * p.local * pe + p.own1 * pe) * es);
                sg.n16 = static_cast<long long>(bytes / 16);
            }
            a.peer_arrived[0] = peer_flags_lo_ + 1;
            a.peer_pushed[0] = peer_flags_lo_ + 3;
            a.my_arrived[0] = d_flags_ + 0;
        }
        if (hi) {
            const SlabSpec p = slab_for_rank(geom_.slices, world_, l.rank + 1, halo_);
            for (int c = 0; c < 3; ++c) {
                PushSeg &sg = a.seg[a.nseg++];
                sg.src = reinterpret_cast<const uint4 *>(plane_ptr(e, buffer, c, s.own1 - n));
                sg.dst = reinterpret_cast<uint4 *>(static_cast<char *>(peer_hi_[idx]) + (c * p.local * pe + (p.own0 - n) * pe) * es);
                sg.n16 = static_cast<long long>(bytes / 16);
            }
            a.peer_arrived[1] = peer_flags_hi_ + 0;
            a.peer_pushed[1] = peer_flags_hi_ + 2;
            a.my_arrived[1] = d_flags_ + 1;
        }
        const int nblk = static_cast<int>(std::max<long long>(1, std::min<long long>(32, (static_cast<long long>(a.nseg) * a.seg[0].n16 + 2047) / 2048)));
        k_push_halo<<<nblk, 256, 0, st>>>(a);
        DCMA_CUDA(cudaGetLastError());
        if (lo) wait_flag(st, d_flags_ + 2, id);
        if (hi) wait_flag(st, d_flags_ + 3, id);
        if (t0) {
            DCMA_CUDA(cudaEventRecord(t1, st));
            x_events_.emplace_back(t0, t1);
        }
        halo_bytes_ += 3 * bytes * ((lo ? 1 : 0) + (hi ? 1 : 0));
        ++exchanges_;
        return;
    }
    k_raise_flags<<<1, 32, 0, st>>>(lo ? peer_flags_lo_ + 1 : nullptr, hi ? peer_flags_hi_ + 0 : nullptr, id);
    DCMA_CUDA(cudaGetLastError());
    if (lo) wait_flag(st, d_flags_ + 0, id);
    if (hi) wait_flag(st, d_flags_ + 1, id);
    if (lo) {  // my first owned planes -> the lower neighbour's upper halo
        const SlabSpec p = slab_for_rank(geom_.slices, world_, l.rank - 1, halo_);
        for (int c = 0; c < 3; ++c)
            DCMA_CUDA(cudaMemcpyAsync(static_cast<char *>(peer_lo_[idx]) + (c * p.local * pe + p.own1 * pe) * es,
                                      plane_ptr(e, buffer, c, s.own0), bytes, cudaMemcpyDefault, st));
    }
    if (hi) {  // my last owned planes -> the upper neighbour's lower halo
        const SlabSpec p = slab_for_rank(geom_.slices, world_, l.rank + 1, halo_);
        for (int c = 0; c < 3; ++c)
            DCMA_CUDA(cudaMemcpyAsync(static_cast<char *>(peer_hi_[idx]) + (c * p.local * pe + (p.own0 - n) * pe) * es,
                                      plane_ptr(e, buffer, c, s.own1 - n), bytes, cudaMemcpyDefault, st));
    }
    k_raise_flags<<<1, 32, 0, st>>>(lo ? peer_flags_lo_ + 3 : nullptr, hi ? peer_flags_hi_ + 2 : nullptr, id);
    DCMA_CUDA(cudaGetLastError());
    if (lo) wait_flag(st, d_flags_ + 2, id);
    if (hi) wait_flag(st, d_flags_ + 3, id);
    if (t0) {
        DCMA_CUDA(cudaEventRecord(t1, st));
        x_events_.emplace_back(t0, t1);
    }
    halo_bytes_ += 3 * bytes * ((lo ? 1 : 0) + (hi ? 1 : 0));
    ++exchanges_;
}

// A 4-byte send/recv with each z-neighbour in stream order: when it completes on this stream, the neighbours' streams
// have reached the same point (their kernels enqueued before it are done as far as their own stream order goes).
void SlabGroup::token_sync() {
    Local &l = *locals_[0];
    cudaStream_t st = static_cast<cudaStream_t>(l.eng->stream_handle());
    DCMA_NCCL(nccl().GroupStart());
    if (l.rank > 0) {
        DCMA_NCCL(nccl().Send(d_token_ + 0, 1, ncclInt32, l.rank - 1, static_cast<ncclComm_t>(comm_), st));
        DCMA_NCCL(nccl().Recv(d_token_ + 1, 1, ncclInt32, l.rank - 1, static_cast<ncclComm_t>(comm_), st));
    }
    if (l.rank + 1 < world_) {
        DCMA_NCCL(nccl().Send(d_token_ + 2, 1, ncclInt32, l.rank + 1, static_cast<ncclComm_t>(comm_), st));
        DCMA_NCCL(nccl().Recv(d_token_ + 3, 1, ncclInt32, l.rank + 1, static_cast<ncclComm_t>(comm_), st));
    }
    DCMA_NCCL(nccl().GroupEnd());
}

void SlabGroup::load(const float *fixed, const float *moving, bool on_device) {
    // One rank per process, host volumes: every rank holds the whole moving image on ITS host, but pushing it whole through
    // every rank's PCIe link (N x the volume through one host) is what bounds a sharded call end to end.  Each rank uploads
    // 1/N of the moving image and the ranks replicate it among themselves over NVLink (grouped ncclSend/ncclRecv: an
    // all-gather of contiguous, possibly unequal slice ranges).  DCMA_B200_SLAB_REPLICATE_VIA_PCIE=1 keeps the plain upload.
    const char *plain = std::getenv("DCMA_B200_SLAB_REPLICATE_VIA_PCIE");
    if (comm_ && world_ > 1 && !on_device && !(plain && *plain == '1')) {
        Local &l = *locals_[0];
        EngineBase *e = l.eng.get();
        const SlabSpec mine = slab_for_rank(geom_.slices, world_, l.rank, 0);
        e->load_partial_moving(fixed, moving, mine.zoff, mine.zoff + mine.local);
        DCMA_CUDA(cudaSetDevice(l.device));
        cudaStream_t st = static_cast<cudaStream_t>(e->stream_handle());
        const size_t plane_bytes = sizeof(float) * static_cast<size_t>(geom_.rows) * geom_.cols * geom_.channels;
        char *base = reinterpret_cast<char *>(e->moving_dev_mutable());
        DCMA_NCCL(nccl().GroupStart());
        for (int p = 0; p < world_; ++p) {
            if (p == l.rank) continue;
            const SlabSpec theirs = slab_for_rank(geom_.slices, world_, p, 0);
            DCMA_NCCL(nccl().Send(base + static_cast<size_t>(mine.zoff) * plane_bytes, static_cast<size_t>(mine.local) * plane_bytes, ncclChar, p,
                                  static_cast<ncclComm_t>(comm_), st));
            DCMA_NCCL(nccl().Recv(base + static_cast<size_t>(theirs.zoff) * plane_bytes, static_cast<size_t>(theirs.local) * plane_bytes, ncclChar, p,
                                  static_cast<ncclComm_t>(comm_), st));
        }
        DCMA_NCCL(nccl().GroupEnd());
        DCMA_CUDA(cudaStreamSynchronize(st));
        return;
    }
    for (auto &l : locals_) l->eng->load(fixed, moving, on_device);
}

// planes [l0, l0 + n) of component c of `buffer`
static inline char *plane_ptr(EngineBase *e, int buffer, int c, int64_t l0) {
    return static_cast<char *>(e->field_base(buffer)) + (c * e->field_comp_stride() + l0 * e->plane_elems()) * e->field_elem_size(buffer);
}

void SlabGroup::exchange(int buffer, int64_t n) {
    if (world_ == 1 || n <= 0) return;
    if (n > halo_) throw std::logic_error("exchange wider than the halo");
    if (comm_ && flags_ok_) {
        exchange_flags(buffer, n, static_cast<cudaStream_t>(locals_[0]->eng->stream_handle()));
        return;
    }
    if (comm_ && ipc_ok_) {
        // ---- NVLink copy engines: PULL the neighbours' boundary planes out of their (IPC-mapped) field buffers into my
        // halo.  The tokens order the streams: before -- the neighbours have produced the planes; after -- they may
        // overwrite them again (nobody is still reading).
        Local &l = *locals_[0];
        EngineBase *e = l.eng.get();
        const SlabSpec s = e->slab();
        const int64_t pe = e->plane_elems(), es = e->field_elem_size(buffer);
        const size_t bytes = static_cast<size_t>(n) * pe * es;
        cudaStream_t st = static_cast<cudaStream_t>(e->stream_handle());
        DCMA_CUDA(cudaSetDevice(l.device));
        int idx = -1;  // which of the field allocations currently plays `buffer` (the roles rotate identically on all ranks)
        for (int i = 0; i < e->field_allocation_count(); ++i)
            if (e->field_allocation(i) == e->field_base(buffer)) idx = i;
        if (idx < 0) throw std::logic_error("field buffer is not one of the engine's allocations");
        token_sync();
        if (l.rank > 0) {
            const SlabSpec p = slab_for_rank(geom_.slices, world_, l.rank - 1, halo_);
            for (int c = 0; c < 3; ++c)
                DCMA_CUDA(cudaMemcpyAsync(plane_ptr(e, buffer, c, s.own0 - n),
                                          static_cast<const char *>(peer_lo_[idx]) + (c * p.local * pe + (p.own1 - n) * pe) * es, bytes,
                                          cudaMemcpyDefault, st));
        }
        if (l.rank + 1 < world_) {
            const SlabSpec p = slab_for_rank(geom_.slices, world_, l.rank + 1, halo_);
            for (int c = 0; c < 3; ++c)
                DCMA_CUDA(cudaMemcpyAsync(plane_ptr(e, buffer, c, s.own1),
                                          static_cast<const char *>(peer_hi_[idx]) + (c * p.local * pe + p.own0 * pe) * es, bytes,
                                          cudaMemcpyDefault, st));
        }
        token_sync();
        halo_bytes_ += 3 * bytes * ((l.rank > 0 ? 1 : 0) + (l.rank + 1 < world_ ? 1 : 0));
        return;
    }
    if (comm_) {
        // ---- NCCL: one grouped send/recv set with both z-neighbours, on the engine's stream
        Local &l = *locals_[0];
        EngineBase *e = l.eng.get();
        const SlabSpec s = e->slab();
        const size_t bytes = static_cast<size_t>(n) * e->plane_elems() * e->field_elem_size(buffer);
        cudaStream_t st = static_cast<cudaStream_t>(e->stream_handle());
        DCMA_CUDA(cudaSetDevice(l.device));
        DCMA_NCCL(nccl().GroupStart());
        for (int c = 0; c < 3; ++c) {
            if (l.rank > 0) {  // lower neighbour: my first owned planes go down, its last owned planes come into my lower halo
                DCMA_NCCL(nccl().Send(plane_ptr(e, buffer, c, s.own0), bytes, ncclChar, l.rank - 1, static_cast<ncclComm_t>(comm_), st));
                DCMA_NCCL(nccl().Recv(plane_ptr(e, buffer, c, s.own0 - n), bytes, ncclChar, l.rank - 1, static_cast<ncclComm_t>(comm_), st));
            }
            if (l.rank + 1 < world_) {
                DCMA_NCCL(nccl().Send(plane_ptr(e, buffer, c, s.own1 - n), bytes, ncclChar, l.rank + 1, static_cast<ncclComm_t>(comm_), st));
                DCMA_NCCL(nccl().Recv(plane_ptr(e, buffer, c, s.own1), bytes, ncclChar, l.rank + 1, static_cast<ncclComm_t>(comm_), st));
            }
        }
        DCMA_NCCL(nccl().GroupEnd());
        halo_bytes_ += 2 * 3 * bytes * ((l.rank > 0 ? 1 : 0) + (l.rank + 1 < world_ ? 1 : 0));
        return;
    }
    // ---- one process: peer copies, ordered by events (producer ready -> copy on the receiver's stream -> both sides
    // wait for the copies before they touch the planes again)
    for (auto &l : locals_) {
        DCMA_CUDA(cudaSetDevice(l->device));
        DCMA_CUDA(cudaEventRecord(l->ev_ready, static_cast<cudaStream_t>(l->eng->stream_handle())));
    }
    for (int r = 0; r + 1 < world_; ++r) {
        Local &lo = *locals_[r], &hi = *locals_[r + 1];
        EngineBase *a = lo.eng.get(), *b = hi.eng.get();
        const SlabSpec sa = a->slab(), sb = b->slab();
        const size_t bytes = static_cast<size_t>(n) * a->plane_elems() * a->field_elem_size(buffer);
        cudaStream_t st_a = static_cast<cudaStream_t>(a->stream_handle()), st_b = static_cast<cudaStream_t>(b->stream_handle());
        // up: a's last owned planes -> b's lower halo, on b's stream
        DCMA_CUDA(cudaSetDevice(hi.device));
        DCMA_CUDA(cudaStreamWaitEvent(st_b, lo.ev_ready, 0));
        for (int c = 0; c < 3; ++c)
            DCMA_CUDA(cudaMemcpyPeerAsync(plane_ptr(b, buffer, c, sb.own0 - n), hi.device, plane_ptr(a, buffer, c, sa.own1 - n), lo.device, bytes, st_b));
        // down: b's first owned planes -> a's upper halo, on a's stream
        DCMA_CUDA(cudaSetDevice(lo.device));
        DCMA_CUDA(cudaStreamWaitEvent(st_a, hi.ev_ready, 0));
        for (int c = 0; c < 3; ++c)
            DCMA_CUDA(cudaMemcpyPeerAsync(plane_ptr(a, buffer, c, sa.own1), lo.device, plane_ptr(b, buffer, c, sb.own0), hi.device, bytes, st_a));
        halo_bytes_ += 2 * 3 * bytes;
    }
    for (auto &l : locals_) {
        DCMA_CUDA(cudaSetDevice(l->device));
        DCMA_CUDA(cudaEventRecord(l->ev_done, static_cast<cudaStream_t>(l->eng->stream_handle())));
    }
    for (int r = 0; r + 1 < world_; ++r) {
        Local &lo = *locals_[r], &hi = *locals_[r + 1];
        DCMA_CUDA(cudaSetDevice(lo.device));
        DCMA_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(lo.eng->stream_handle()), hi.ev_done, 0));
        DCMA_CUDA(cudaSetDevice(hi.device));
        DCMA_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(hi.eng->stream_handle()), lo.ev_done, 0));
    }
}

void SlabGroup::gather_mse_and_stop_rule(long long it) {
    for (auto &l : locals_) l->eng->publish_mse(l->d_mse2);
    if (comm_) {
        Local &l = *locals_[0];
        DCMA_CUDA(cudaSetDevice(l.device));
        DCMA_NCCL(nccl().AllGather(l.d_mse2, l.d_gather, 2, ncclDouble, static_cast<ncclComm_t>(comm_),
                                   static_cast<cudaStream_t>(l.eng->stream_handle())));
    } else {
        for (auto &l : locals_) {
            DCMA_CUDA(cudaSetDevice(l->device));
            DCMA_CUDA(cudaEventRecord(l->ev_ready, static_cast<cudaStream_t>(l->eng->stream_handle())));
        }
        for (auto &dst : locals_) {
            DCMA_CUDA(cudaSetDevice(dst->device));
            cudaStream_t st = static_cast<cudaStream_t>(dst->eng->stream_handle());
            for (auto &src : locals_) {
                if (src.get() != dst.get()) DCMA_CUDA(cudaStreamWaitEvent(st, src->ev_ready, 0));
                DCMA_CUDA(cudaMemcpyPeerAsync(dst->d_gather + 2 * src->rank, dst->device, src->d_mse2, src->device, 2 * sizeof(double), st));
            }
        }
        // d_mse2 is rewritten only by the next iteration's publish, which every stream reaches after the exchanges
        // below have synchronised it with its neighbours; a full fence keeps this independent of the stage pattern
        for (auto &l : locals_) {
            DCMA_CUDA(cudaSetDevice(l->device));
            DCMA_CUDA(cudaEventRecord(l->ev_done, static_cast<cudaStream_t>(l->eng->stream_handle())));
        }
        for (auto &a : locals_)
            for (auto &b : locals_)
                if (a.get() != b.get()) {
                    DCMA_CUDA(cudaSetDevice(a->device));
                    DCMA_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(a->eng->stream_handle()), b->ev_done, 0));
                }
    }
    for (auto &l : locals_) l->eng->stop_rule(l->d_gather, world_, it);
}

void SlabGroup::run(int64_t max_iterations, double *mse_history_host, dcma_b200_run_stats *stats) {
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (max_iterations <= 0) return;
    halo_bytes_ = 0;
    EngineBase *e0 = locals_[0]->eng.get();
    const bool diffeo = e0->diffeomorphic();
    const int64_t rz_u = e0->update_smoothed() ? e0->smoothing_radius_z(true) : 0;
    const int64_t rz_d = e0->deformation_smoothed() ? e0->smoothing_radius_z(false) : 0;
    if (world_ > 1 && std::max(rz_u, rz_d) > halo_) throw std::invalid_argument("halo thinner than the smoothing radius along z");
    const bool can_converge = (params_.convergence_threshold > 0.0);
    const int64_t poll = (params_.check_every > 0) ? params_.check_every : 4;
    for (auto &l : locals_) l->eng->begin_run(max_iterations);
    // With threshold <= 0 the stop rule can never fire (|prev - mse| < threshold is false), so the slabs' MSE terms are
    // only needed for the history: they are kept per iteration on each device and combined once, after the loop,
    // instead of one all-gather per iteration.
    if (!can_converge) {
        for (auto &l : locals_) {
            DCMA_CUDA(cudaSetDevice(l->device));
            cudaFree(l->d_mse_hist);
            l->d_mse_hist = nullptr;
            DCMA_CUDA(cudaMalloc(&l->d_mse_hist, 2 * sizeof(double) * max_iterations));
        }
    }
    // Overlapped schedule (one rank per process, flag transport, diffeomorphic): the per-voxel stages are split so that
    // the planes a neighbour needs are produced first and travel (copy engines, second stream) under the rest:
    //   FORCE   : boundary planes, then  [push U, r_z planes]  ||  interior
    //   SMOOTH U: whole (needs the U halo)
    //   COMPOSE : [push U~, halo planes] || interior part 1;  boundary planes (need the U~ halo);
    //             [push D, r_z planes]   || interior part 2
    //   SMOOTH D: whole (needs the D halo)
    // Splitting a per-voxel kernel by z-range costs nothing; the smoothing kernels are not split (each extra start of a
    // z march would re-stage 2 r_z slices).
    const char *no_ov = std::getenv("DCMA_B200_SLAB_NO_OVERLAP");
    EngineBase *e = e0;
    const SlabSpec sp = e->slab();
    const int own0 = static_cast<int>(sp.own0), own1 = static_cast<int>(sp.own1), H = static_cast<int>(halo_);
    const bool overlap = comm_ && flags_ok_ && world_ > 1 && diffeo && !e->symmetric() && !(no_ov && *no_ov == '1') &&
                         (own1 - own0) >= 4 * H + 4 && rz_u <= H && rz_d <= H;
    overlapped_ = overlap;
    exchanges_ = 0;
    for (auto &pr : x_events_) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    x_events_.clear();
    cudaStream_t s_main = static_cast<cudaStream_t>(e->stream_handle());
    for (int64_t it = 0; it < max_iterations; ++it) {
        if (overlap) {
            DCMA_CUDA(cudaSetDevice(locals_[0]->device));
            const int ru = static_cast<int>(rz_u), rd = static_cast<int>(rz_d);
            // ---- FORCE
            if (ru > 0) {
                e->force_range(it, own0, own0 + ru);
                e->force_range(it, own1 - ru, own1);
                DCMA_CUDA(cudaEventRecord(ev_main_, s_main));
                DCMA_CUDA(cudaStreamWaitEvent(comm_stream_, ev_main_, 0));
                exchange_flags(DCMA_B200_BUF_UPDATE, ru, comm_stream_);
                DCMA_CUDA(cudaEventRecord(ev_x_[0], comm_stream_));
                e->force_range(it, own0 + ru, own1 - ru);
                e->force_finish(it);
            } else {
                e->iter_stage(IterStage::Force, it, true);
            }
            if (can_converge) {
                gather_mse_and_stop_rule(it);
            } else {
                e->publish_mse(locals_[0]->d_mse_hist + 2 * it);
                e->stop_rule(locals_[0]->d_mse_hist + 2 * it, 1, it);
            }
            if (ru > 0) {
                DCMA_CUDA(cudaStreamWaitEvent(s_main, ev_x_[0], 0));
                e->iter_stage(IterStage::SmoothU, it, true);
            }
            // ---- COMPOSE
            DCMA_CUDA(cudaEventRecord(ev_main_, s_main));
            DCMA_CUDA(cudaStreamWaitEvent(comm_stream_, ev_main_, 0));
            exchange_flags(DCMA_B200_BUF_UPDATE, H, comm_stream_);
            DCMA_CUDA(cudaEventRecord(ev_x_[1], comm_stream_));
            const int mid = (own0 + own1) / 2;
            e->update_range(it, own0 + H, mid);
            DCMA_CUDA(cudaStreamWaitEvent(s_main, ev_x_[1], 0));
            e->update_range(it, own0, own0 + H);
            e->update_range(it, own1 - H, own1);
            if (rd > 0) {
                DCMA_CUDA(cudaEventRecord(ev_main_, s_main));
                DCMA_CUDA(cudaStreamWaitEvent(comm_stream_, ev_main_, 0));
                exchange_flags(DCMA_B200_BUF_DEFORMATION, rd, comm_stream_);
                DCMA_CUDA(cudaEventRecord(ev_x_[2], comm_stream_));
            }
            e->update_range(it, mid, own1 - H);
            if (rd > 0) DCMA_CUDA(cudaStreamWaitEvent(s_main, ev_x_[2], 0));
            e->iter_stage(IterStage::SmoothD, it, true);
            if (can_converge && ((it + 1) % poll == 0) && (it + 1 < max_iterations) && e0->converged_poll()) break;
            continue;
        }
        for (auto &l : locals_) l->eng->iter_stage(IterStage::Force, it, /*global_mse=*/true);
        if (can_converge) {
            gather_mse_and_stop_rule(it);
        } else {
            for (auto &l : locals_) {
                l->eng->publish_mse(l->d_mse_hist + 2 * it);
                l->eng->stop_rule(l->d_mse_hist + 2 * it, 1, it);  // counts the iteration; history is rebuilt below
            }
        }
        if (rz_u > 0) {
            exchange(DCMA_B200_BUF_UPDATE, rz_u);
            for (auto &l : locals_) l->eng->iter_stage(IterStage::SmoothU, it, true);
        }
        if (diffeo) exchange(DCMA_B200_BUF_UPDATE, halo_);  // the composition gathers U~ at p + D(p)
        for (auto &l : locals_) l->eng->iter_stage(IterStage::Update, it, true);
        if (rz_d > 0) exchange(DCMA_B200_BUF_DEFORMATION, rz_d);
        for (auto &l : locals_) l->eng->iter_stage(IterStage::SmoothD, it, true);
        // symmetric force (extension a11): the next force stage needs W = warp(M, D) one plane beyond the owned range to
        // take the z gradient of W at the slab faces, i.e. one plane of the FINAL D of this iteration from each neighbour
        if (e0->symmetric() && it + 1 < max_iterations) exchange(DCMA_B200_BUF_DEFORMATION, 1);
        if (can_converge && ((it + 1) % poll == 0) && (it + 1 < max_iterations) && e0->converged_poll()) break;
    }
    if (comm_stream_) DCMA_CUDA(cudaStreamSynchronize(comm_stream_));
    dcma_b200_run_stats st0;
    bool first = true;
    for (auto &l : locals_) {
        dcma_b200_run_stats st;
        l->eng->end_run(first ? mse_history_host : nullptr, &st);
        if (first) {
            st0 = st;
        } else {
            st0.loop_ms = std::max(st0.loop_ms, st.loop_ms);
            st0.kernel_launches += st.kernel_launches;
        }
        first = false;
    }
    exchange_ms_ = 0.0;
    for (auto &pr : x_events_) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) exchange_ms_ += t;
    }
    if (x_events_.size() < static_cast<size_t>(exchanges_) && !x_events_.empty())
        exchange_ms_ *= static_cast<double>(exchanges_) / static_cast<double>(x_events_.size());  // sampled
    if (stats) *stats = st0;
    if (!can_converge && mse_history_host && st0.iterations_done > 0) {
        // global MSE history: sum the slabs' terms in rank order (the order of k_stop_rule)
        const int64_t n = st0.iterations_done;
        std::vector<double> all(static_cast<size_t>(2 * n) * world_, 0.0);
        if (comm_) {
            Local &l = *locals_[0];
            DCMA_CUDA(cudaSetDevice(l.device));
            double *d_all = nullptr;
            DCMA_CUDA(cudaMalloc(&d_all, sizeof(double) * all.size()));
            cudaStream_t st = static_cast<cudaStream_t>(l.eng->stream_handle());
            DCMA_NCCL(nccl().AllGather(l.d_mse_hist, d_all, 2 * n, ncclDouble, static_cast<ncclComm_t>(comm_), st));
            DCMA_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(double) * all.size(), cudaMemcpyDeviceToHost, st));
            DCMA_CUDA(cudaStreamSynchronize(st));
            cudaFree(d_all);
        } else {
            for (auto &l : locals_) {
                DCMA_CUDA(cudaSetDevice(l->device));
                DCMA_CUDA(cudaMemcpy(all.data() + static_cast<size_t>(2 * n) * l->rank, l->d_mse_hist, sizeof(double) * 2 * n,
                                     cudaMemcpyDeviceToHost));
            }
        }
        for (int64_t i = 0; i < n; ++i) {
            double sum = 0.0, cnt = 0.0;
            for (int r = 0; r < world_; ++r) {
                sum += all[static_cast<size_t>(2 * n) * r + 2 * i];
                cnt += all[static_cast<size_t>(2 * n) * r + 2 * i + 1];
            }
            mse_history_host[i] = (cnt > 0.0) ? sum / cnt : sum;
        }
    }
    for (auto &l : locals_)
        if (l->eng->halo_overflowed())
            throw std::logic_error("a composition gather left the slab halo (|D_z| too large for halo_planes = " +
                                   std::to_string(halo_) + "); re-run with a larger halo or fewer slabs");
}

void SlabGroup::owned_range(int64_t *z0, int64_t *z1) const {
    // the slices owned by THIS process (all of them for a one-process group)
    const SlabSpec a = locals_.front()->eng->slab(), b = locals_.back()->eng->slab();
    *z0 = a.zoff + a.own0;
    *z1 = b.zoff + b.own1;
}

void SlabGroup::get_deformation(double *host_aos) {
    int64_t z0, z1;
    owned_range(&z0, &z1);
    const int64_t plane3 = 3 * geom_.rows * geom_.cols;
    for (auto &l : locals_) {
        const SlabSpec s = l->eng->slab();
        l->eng->get_owned_deformation(host_aos + (s.zoff + s.own0 - z0) * plane3);
    }
}

void SlabGroup::import_deformation_device(const double *device_aos_whole) {
    for (auto &l : locals_) l->eng->import_deformation_whole(device_aos_whole);
    for (auto &l : locals_) l->eng->sync();
}

int64_t SlabGroup::device_bytes() const {
    int64_t b = 0;
    for (auto &l : locals_) b += l->eng->device_bytes();
    return b;
}

}  // namespace dcma
