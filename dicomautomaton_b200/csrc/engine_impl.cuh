// dcma_b200 -- Engine<T, STRICT, TU>: owns the HBM buffers and launches the kernels of one registration.  T is the storage
// and arithmetic type of the deformation field D, TU the storage type of the update field U (== T except in
// DCMA_B200_FIELD_MIXED, where D stays double and U is stored and smoothed as float).
// Included by exactly four translation units (engine_f64.cu, engine_f64_strict.cu, engine_f32.cu, engine_mixed.cu), each with its own
// DCMA_NS so that no device function compiled with different flags (-fmad=false for STRICT) can be shared by accident.
//
// Mirrors the reference engine (src/Alignment_Demons.cc:52-704):
//   ctor                      -> cc:54-75, 152-187 (buffers; gradient is recomputed on the fly, never stored)
//   load()                    -> cc:181-186
//   run()                     -> AlignViaDemons' loop cc:980-995 over compute_single_iteration cc:81-108
//   stage()                   -> the engine's private methods, one at a time (stage-level parity tests)
//   get(DEFORMATION)          -> export_deformation_volume cc:110-113
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <tuple>
#include <type_traits>
#include <vector>

#include "demons_kernels.cuh"
#include "device_pool.h"
#include "engine.h"
#include "smooth_kernels.cuh"
#include "smooth_tma.cuh"

#ifndef DCMA_SMOOTH_TMA
#define DCMA_SMOOTH_TMA 1  // fp64 fields: k_smooth3d_tma (TMA-staged, wide x-pass tasks) whenever its layout rules hold
#endif
#ifndef DCMA_FUSED_REDUCE
#define DCMA_FUSED_REDUCE 1  // 1: the force kernel's last block reduces the MSE partials; 0: k_reduce_final as its own launch
#endif

namespace DCMA_NS {
using namespace ::dcma;

template <typename T>
struct SmoothPlan {
    bool active = false;  // sigma > 0 (cc:564-566)
    bool fast = false;    // fused k_smooth3d usable (every radius <= kMaxFastRadius)
    int r[3] = {0, 0, 0};
    int rk = 0;                // radius the fused kernel is instantiated for: max of r[]
    std::vector<double> w[3];  // normalised weights per axis (cc:587-597)
    SmoothWeights<T> W;
    T *tab[3] = {nullptr, nullptr, nullptr};        // per-position wsum (STRICT) or 1/wsum
    double *dw[3] = {nullptr, nullptr, nullptr};    // weights in global memory for k_smooth_axis
};

template <typename T, bool STRICT, typename TU = T>
class Engine final : public EngineBase {
    static constexpr bool kSameU = std::is_same<T, TU>::value;

  public:
    Engine(const dcma_b200_geom &geom, const dcma_b200_demons_params &params, void *stream_in, const SlabSpec &slab_in)
        : p_(params), slab_(slab_in) {
        if (geom.slices < 1 || geom.rows < 1 || geom.cols < 1 || geom.channels < 1)
            throw std::invalid_argument("volume extents and channel count must be >= 1");
        if (!(geom.pxl_dx > 0.0) || !(geom.pxl_dy > 0.0) || !(geom.pxl_dz > 0.0))
            throw std::invalid_argument("voxel spacing must be > 0");
        if (geom.slices > (1 << 19) || geom.rows > (1 << 19) || geom.cols > (1 << 19))
            throw std::invalid_argument("volume extent too large");
        if (slab_.local <= 0) slab_ = SlabSpec{0, geom.slices, 0, geom.slices};  // whole volume
        if (slab_.zoff < 0 || slab_.zoff + slab_.local > geom.slices || slab_.own0 < 0 || slab_.own1 > slab_.local ||
            slab_.own0 >= slab_.own1)
            throw std::invalid_argument("bad slab specification");
        const long long Ng = geom.slices * geom.rows * geom.cols;     // voxels of the whole volume (moving image)
        const long long N = slab_.local * geom.rows * geom.cols;      // voxels this engine holds
        if (Ng * 3 >= (1LL << 40)) throw std::invalid_argument("volume too large");
        g_.S = static_cast<int>(slab_.local);
        g_.Sg = static_cast<int>(geom.slices);
        g_.zoff = static_cast<int>(slab_.zoff);
        g_.zb = static_cast<int>(slab_.own0);
        g_.ze = static_cast<int>(slab_.own1);
        g_.R = static_cast<int>(geom.rows);
        g_.C = static_cast<int>(geom.cols);
        g_.CH = static_cast<int>(geom.channels);
        g_.N = N;
        Ng_ = Ng;
        g_.px = geom.pxl_dx;
        g_.py = geom.pxl_dy;
        g_.pz = geom.pxl_dz;
        g_.ipx = 1.0 / geom.pxl_dx;
        g_.ipy = 1.0 / geom.pxl_dy;
        g_.ipz = 1.0 / geom.pxl_dz;
        g_.is2d = (geom.slices == 1) ? 1 : 0;
        small_index_ = (Ng * geom.channels < 2000000000LL);  // 32-bit element offsets inside the per-voxel kernels

        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw CudaError(cudaErrorNoDevice, "no CUDA device is available (dcma_b200 has no CPU fallback)");
        if (p_.device >= 0) {
            DCMA_CUDA(cudaSetDevice(p_.device));
            dev_ = p_.device;
        } else {
            DCMA_CUDA(cudaGetDevice(&dev_));
        }
        int cc_major = 0, cc_minor = 0;
        char dev_name[256] = {0};
        DCMA_CUDA(cached_device_info(dev_, &cc_major, &cc_minor, &sms_, dev_name));
        if (cc_major < 10)
            throw CudaError(cudaErrorInvalidDevice, std::string("dcma_b200 kernels are built for sm_100a only; device is ") +
                                                        dev_name + " (sm_" + std::to_string(cc_major) +
                                                        std::to_string(cc_minor) + ")");
        if (stream_in) {
            stream_ = static_cast<cudaStream_t>(stream_in);
        } else {
            DCMA_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
            own_stream_ = true;
        }
        // from here on a throwing allocation must not leak what the constructor already holds: the destructor does not
        // run for a half-built object
        try {
            allocate_and_plan(geom, Ng, N);
        } catch (...) {
            release_all();
            throw;
        }
    }

  private:
    void allocate_and_plan(const dcma_b200_geom &geom, const long long Ng, const long long N) {
        // per-voxel kernels: 3-D grid of 32x8 tiles, striding over the owned slices; ~16 resident-CTA waves of blocks
        const int ntx = (g_.C + kTileX - 1) / kTileX, nty = (g_.R + kTileY - 1) / kTileY;
        if (nty > 65535) throw std::invalid_argument("too many rows");
        const long long per_plane = static_cast<long long>(ntx) * nty;
        const int So = g_.ze - g_.zb;  // owned slices
        // runs of ~16 consecutive slices per block (rolling z stencil, gather reuse), at least ~8 waves of blocks
        const long long want = static_cast<long long>(sms_) * 8 * 8;
        int gz = static_cast<int>(std::max<long long>((So + 15) / 16, std::min<long long>(So, (want + per_plane - 1) / per_plane)));
        gz = std::max(1, std::min(gz, std::min(So, 65535)));
        zlen_ = (So + gz - 1) / gz;
        gz = (So + zlen_ - 1) / zlen_;
        grid3_ = dim3(ntx, nty, gz);
        grid_vox_ = static_cast<int>(per_plane * grid3_.z);  // number of MSE partials
        grid_lin_ = static_cast<int>(std::min<long long>((3 * N + 255) / 256, static_cast<long long>(sms_) * 16));

        dF_ = alloc<float>(N * g_.CH);
        dM_ = alloc<float>(Ng * g_.CH);  // the moving image is always held whole: the warp gathers have no a-priori bound
        dD_ = alloc<T>(3 * N);
        dU_ = alloc<TU>(3 * N);
        dTmp_ = alloc<T>(3 * N);
        field_alloc_[0] = dD_;
        field_alloc_[1] = dU_;
        field_alloc_[2] = dTmp_;
        if constexpr (!kSameU) {  // U cannot share D's scratch field: its own, in its own type (same total bytes as 3 x T)
            dTmpU_ = alloc<TU>(3 * N);
            field_alloc_[3] = dTmpU_;
            n_field_alloc_ = 4;
        }
        // slack: a force stage split into up to kMaxForceParts z-ranges rounds each range up to whole z-chunks
        part_cap_ = grid_vox_ + kMaxForceParts * static_cast<int>(per_plane);
        d_part_sum_ = alloc<double>(part_cap_);
        d_part_cnt_ = alloc<long long>(part_cap_);
        d_ctrl_ = alloc<Ctrl>(1);
        static_assert(sizeof(Ctrl) <= PinnedSmallPool::kBlock, "control block larger than a pooled pinned block");
        {
            void *hp = nullptr;
            DCMA_CUDA(PinnedSmallPool::instance().acquire(&hp));
            h_ctrl_ = static_cast<Ctrl *>(hp);
        }

        build_plan(plan_u_, p_.update_field_smoothing_sigma);
        build_plan(plan_d_, p_.deformation_field_smoothing_sigma);
        reset_ctrl();
        DCMA_CUDA(cudaMemsetAsync(dD_, 0, sizeof(T) * 3 * N, stream_));
        DCMA_CUDA(cudaMemsetAsync(dU_, 0, sizeof(TU) * 3 * N, stream_));
    }
    void release_all() noexcept {
        cudaSetDevice(dev_);
        if (stream_) cudaStreamSynchronize(stream_);
        for (auto &q : allocs_) DevicePool::instance().release(dev_, q.first, q.second);
        allocs_.clear();
        PinnedSmallPool::instance().release(h_ctrl_);
        drop_history();
        if (ev_run0_) cudaEventDestroy(ev_run0_);
        if (ev_run1_) cudaEventDestroy(ev_run1_);
        for (cudaEvent_t e : ev_stage_)
            if (e) cudaEventDestroy(e);
        for (auto &pr : prof_) {
            cudaEventDestroy(pr.a);
            cudaEventDestroy(pr.b);
        }
        prof_.clear();
        if (copy_stream_) cudaStreamDestroy(copy_stream_);
        if (own_stream_ && stream_) cudaStreamDestroy(stream_);
        h_ctrl_ = nullptr;
        d_hist_ = nullptr;
        ev_run0_ = ev_run1_ = nullptr;
        ev_stage_[0] = ev_stage_[1] = nullptr;
        copy_stream_ = nullptr;
        stream_ = nullptr;
        cudaGetLastError();  // a failed allocation leaves a sticky-looking (non-fatal) error code behind
    }

  public:
    ~Engine() override { release_all(); }

    int device() const override { return dev_; }
    int64_t device_bytes() const override { return bytes_; }
    void sync() override {
        DCMA_CUDA(cudaSetDevice(dev_));
        DCMA_CUDA(cudaStreamSynchronize(stream_));
    }

    // ---- cc:181-186 ------------------------------------------------------------------------------------------
    // `fixed` and `moving` always point at WHOLE volumes; a slab engine takes its own slices of the fixed image.
    void load(const float *fixed, const float *moving, bool on_device) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        const size_t plane_bytes = sizeof(float) * static_cast<size_t>(g_.R) * g_.C * g_.CH;
        DCMA_CUDA(cudaMemcpyAsync(dF_, fixed + static_cast<size_t>(g_.zoff) * g_.R * g_.C * g_.CH, plane_bytes * g_.S, kind,
                                  stream_));
        DCMA_CUDA(cudaMemcpyAsync(dM_, moving, sizeof(float) * Ng_ * g_.CH, kind, stream_));
        DCMA_CUDA(cudaMemsetAsync(dD_, 0, sizeof(T) * 3 * g_.N, stream_));
        DCMA_CUDA(cudaMemsetAsync(dU_, 0, sizeof(TU) * 3 * g_.N, stream_));
        w_src_ = 0;  // W := M (cc:183)
        reset_ctrl();
        if (!on_device) DCMA_CUDA(cudaStreamSynchronize(stream_));  // host buffers may be released by the caller
        loaded_ = true;
    }

    void load_partial_moving(const float *fixed, const float *moving, int64_t mz0, int64_t mz1) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        if (mz0 < 0 || mz1 > g_.Sg || mz0 > mz1) throw std::invalid_argument("bad moving slice range");
        const size_t plane = static_cast<size_t>(g_.R) * g_.C * g_.CH;
        DCMA_CUDA(cudaMemcpyAsync(dF_, fixed + static_cast<size_t>(g_.zoff) * plane, sizeof(float) * plane * g_.S, cudaMemcpyHostToDevice, stream_));
        DCMA_CUDA(cudaMemcpyAsync(dM_ + static_cast<size_t>(mz0) * plane, moving + static_cast<size_t>(mz0) * plane,
                                  sizeof(float) * plane * static_cast<size_t>(mz1 - mz0), cudaMemcpyHostToDevice, stream_));
        DCMA_CUDA(cudaMemsetAsync(dD_, 0, sizeof(T) * 3 * g_.N, stream_));
        DCMA_CUDA(cudaMemsetAsync(dU_, 0, sizeof(TU) * 3 * g_.N, stream_));
        w_src_ = 0;
        reset_ctrl();
        DCMA_CUDA(cudaStreamSynchronize(stream_));
        loaded_ = true;
    }
    float *moving_dev_mutable() override { return dM_; }

    // ---- AlignViaDemons loop, cc:980-995 -----------------------------------------------------------------------
    void run(int64_t max_iterations, bool profile, double *mse_history_host, dcma_b200_run_stats *stats) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        if (!loaded_) throw std::logic_error("engine_run called before engine_load");
        if (is_slab()) throw std::logic_error("a z-slab engine is driven by its slab group, not by engine_run");
        if (stats) std::memset(stats, 0, sizeof(*stats));
        if (max_iterations <= 0) return;
        begin_run(max_iterations);
        profiling_ = profile;

        // The stop rule runs on the device (tail of the force kernel); the host only polls it.  With threshold <= 0
        // the rule `|prev-mse| < threshold` can never fire, so no poll is needed at all.
        const bool can_converge = (p_.convergence_threshold > 0.0);
        const int64_t poll = (p_.check_every > 0) ? p_.check_every : 4;
        for (int64_t it = 0; it < max_iterations; ++it) {
            iteration(it, d_hist_);
            if (can_converge && ((it + 1) % poll == 0) && (it + 1 < max_iterations) && converged_poll()) break;
        }
        end_run(mse_history_host, stats);
    }

    // ---- the same loop in pieces, for drivers that interleave halo exchanges between the stages (slab.cu) -------
    void begin_run(int64_t max_iterations) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        if (!loaded_) throw std::logic_error("run called before load");
        reset_ctrl();
        drop_history();
        d_hist_bytes_ = sizeof(double) * static_cast<size_t>(std::max<int64_t>(max_iterations, 1));
        {
            void *hp = nullptr;
            DCMA_CUDA(DevicePool::instance().acquire(dev_, d_hist_bytes_, &hp));
            d_hist_ = static_cast<double *>(hp);
        }
        launches_ = 0;
        prof_.clear();
        profiling_ = false;
        rot_hist_.clear();
        if (ev_run0_) cudaEventDestroy(ev_run0_);
        if (ev_run1_) cudaEventDestroy(ev_run1_);
        ev_run0_ = ev_run1_ = nullptr;
        DCMA_CUDA(cudaEventCreate(&ev_run0_));
        DCMA_CUDA(cudaEventCreate(&ev_run1_));
        DCMA_CUDA(cudaEventRecord(ev_run0_, stream_));
    }
    bool converged_poll() override {
        DCMA_CUDA(cudaSetDevice(dev_));
        DCMA_CUDA(cudaMemcpyAsync(h_ctrl_, d_ctrl_, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream_));
        DCMA_CUDA(cudaStreamSynchronize(stream_));
        return h_ctrl_->converged_at != kNotConverged;
    }
    bool halo_overflowed() override {
        DCMA_CUDA(cudaSetDevice(dev_));
        DCMA_CUDA(cudaMemcpyAsync(h_ctrl_, d_ctrl_, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream_));
        DCMA_CUDA(cudaStreamSynchronize(stream_));
        return h_ctrl_->halo_overflow != 0;
    }
    void end_run(double *mse_history_host, dcma_b200_run_stats *stats) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        DCMA_CUDA(cudaEventRecord(ev_run1_, stream_));
        DCMA_CUDA(cudaMemcpyAsync(h_ctrl_, d_ctrl_, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream_));
        DCMA_CUDA(cudaStreamSynchronize(stream_));
        const int64_t done = h_ctrl_->iters_done;
        // Iterations enqueued after the stop rule fired are skipped ON THE DEVICE (every kernel returns early), but the
        // host still rotated the field pointers for them: put the roles back to where they stood at the end of the
        // converging iteration, so that D is the field the reference returns (cc:987-997).
        if (h_ctrl_->converged_at != kNotConverged && h_ctrl_->converged_at >= 0 &&
            static_cast<size_t>(h_ctrl_->converged_at) < rot_hist_.size()) {
            const Rotation &r = rot_hist_[static_cast<size_t>(h_ctrl_->converged_at)];
            dD_ = r.d;
            dU_ = r.u;
            dTmp_ = r.tmp;
            dTmpU_ = r.tmpu;
        }
        rot_hist_.clear();
        if (mse_history_host && done > 0)
            DCMA_CUDA(cudaMemcpy(mse_history_host, d_hist_, sizeof(double) * done, cudaMemcpyDeviceToHost));
        if (p_.verbosity >= 1 && (!is_slab() || g_.zoff + g_.zb == 0)) {
            std::vector<double> h(done);
            if (done > 0) DCMA_CUDA(cudaMemcpy(h.data(), d_hist_, sizeof(double) * done, cudaMemcpyDeviceToHost));
            for (int64_t i = 0; i < done; ++i) std::fprintf(stderr, "Iteration %lld: MSE = %.17g\n", (long long)i, h[i]);
            if (h_ctrl_->converged_at != kNotConverged)
                std::fprintf(stderr, "Converged after %lld iterations\n", (long long)h_ctrl_->converged_at);
        }
        drop_history();
        float ms = 0.f;
        DCMA_CUDA(cudaEventElapsedTime(&ms, ev_run0_, ev_run1_));
        cudaEventDestroy(ev_run0_);
        cudaEventDestroy(ev_run1_);
        ev_run0_ = ev_run1_ = nullptr;
        if (stats) {
            std::memset(stats, 0, sizeof(*stats));
            stats->iterations_done = done;
            stats->kernel_launches = launches_;
            stats->loop_ms = ms;
            for (auto &pr : prof_) {
                float t = 0.f;
                cudaEventElapsedTime(&t, pr.a, pr.b);
                stats->stage_ms[pr.stage] += t;
                stats->stage_launches[pr.stage] += pr.launches;
            }
        }
        for (auto &pr : prof_) {
            cudaEventDestroy(pr.a);
            cudaEventDestroy(pr.b);
        }
        prof_.clear();
        profiling_ = false;
    }
    // one stage of loop iteration `it`; global_mse: the force kernel only publishes this slab's (sum, count), the
    // stop rule is applied by stop_rule() once the slabs' values have been gathered
    void iter_stage(IterStage st, long long it, bool global_mse) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        const bool diffeo = (p_.use_diffeomorphic != 0);
        switch (st) {
            case IterStage::Force:
                if (p_.use_symmetric_force) {
                    // extension a11 on a slab: the gradient of W reads W one plane beyond the owned range, so W is
                    // evaluated on owned +- 1 planes; the slab group has exchanged that plane of the smoothed D
                    materialise_warped(/*z_margin=*/1);
                    launch_force(2, it, d_hist_, it, global_mse ? 0 : 1, /*reduce=*/false, /*fused_reduce=*/true);
                } else {
                    launch_force(w_src_, it, d_hist_, it, global_mse ? 0 : 1, /*reduce=*/false, /*fused_reduce=*/true);
                }
                break;
            case IterStage::SmoothU:
                if (diffeo && plan_u_.active) smooth(dU_, tmp_u(), plan_u_, it);
                break;
            case IterStage::Update:
                if (diffeo)
                    launch_compose(it);
                else
                    launch_add(it);
                break;
            case IterStage::SmoothD:
                if (plan_d_.active) smooth(dD_, dTmp_, plan_d_, it);
                w_src_ = 1;  // the next force kernel warps M through the new D (cc:106)
                record_rotation(it);
                break;
        }
    }
    // ---- the per-voxel stages on a sub-range of the owned slices (slab.cu overlaps the halo exchange of the boundary
    // planes with the interior).  FORCE: any number of force_range calls (disjoint ranges covering the owned slices),
    // then force_finish, which reduces their MSE partials.  The symmetric force reads W with its neighbours and is
    // not split.
    void force_range(long long it, int z0, int z1) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        force_parts_ += launch_force(w_src_, it, d_hist_, it, 0, /*reduce=*/false, /*fused_reduce=*/false, z0, z1, force_parts_);
    }
    void force_finish(long long it) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        launch_reduce(it, d_hist_, it, /*count_iter=*/0, force_parts_);
        force_parts_ = 0;
    }
    void update_range(long long it, int z0, int z1) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        if (p_.use_diffeomorphic != 0)
            launch_compose(it, z0, z1);
        else
            throw std::logic_error("update_range is for the diffeomorphic composition");
    }
    bool symmetric() const override { return p_.use_symmetric_force != 0; }
    void publish_mse(double *dev_out2) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        k_publish_mse<<<1, 32, 0, stream_>>>(d_ctrl_, dev_out2);
        DCMA_CUDA(cudaGetLastError());
        ++launches_;
    }
    void stop_rule(const double *dev_gather, int world, long long it) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        k_stop_rule<<<1, 32, 0, stream_>>>(d_ctrl_, dev_gather, world, d_hist_, it, it);
        DCMA_CUDA(cudaGetLastError());
        ++launches_;
    }
    bool update_smoothed() const override { return p_.use_diffeomorphic != 0 && plan_u_.active; }
    bool deformation_smoothed() const override { return plan_d_.active; }
    bool diffeomorphic() const override { return p_.use_diffeomorphic != 0; }
    int smoothing_radius_z(bool update_field) const override { return update_field ? plan_u_.r[2] : plan_d_.r[2]; }
    void *field_base(int buffer) override { return buffer == DCMA_B200_BUF_UPDATE ? static_cast<void *>(dU_) : static_cast<void *>(dD_); }
    int field_allocation_count() const override { return n_field_alloc_; }
    void *field_allocation(int i) override { return field_alloc_[i]; }
    int64_t field_elem_size(int buffer) const override {
        return static_cast<int64_t>(buffer == DCMA_B200_BUF_UPDATE ? sizeof(TU) : sizeof(T));
    }
    int64_t field_comp_stride() const override { return g_.N; }
    int64_t plane_elems() const override { return static_cast<int64_t>(g_.R) * g_.C; }
    void *stream_handle() override { return stream_; }
    SlabSpec slab() const override { return slab_; }
    void get_owned_deformation(double *host_aos) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        export_field_host(dD_, host_aos);
    }

    // ---- one reference engine method at a time ------------------------------------------------------------------
    void stage(int stage, double sigma_mm, double *mse_out, int64_t *n_valid_out) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        if (!loaded_) throw std::logic_error("engine_stage called before engine_load");
        const long long it = 0;  // stage calls are never skipped: converged_at is reset below
        reset_ctrl();
        switch (stage) {
            case DCMA_B200_STAGE_FORCE: {
                materialise_warped();
                launch_force(2, it, nullptr, 0, 0);
                DCMA_CUDA(cudaMemcpyAsync(h_ctrl_, d_ctrl_, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream_));
                DCMA_CUDA(cudaStreamSynchronize(stream_));
                if (mse_out) *mse_out = h_ctrl_->last_mse;
                if (n_valid_out) *n_valid_out = h_ctrl_->last_n;
                break;
            }
            case DCMA_B200_STAGE_SMOOTH_U:
            case DCMA_B200_STAGE_SMOOTH_D: {
                if (stage == DCMA_B200_STAGE_SMOOTH_U)
                    smooth(dU_, tmp_u(), plan_for<TU>(sigma_mm), it);
                else
                    smooth(dD_, dTmp_, plan_for<T>(sigma_mm), it);
                break;
            }
            case DCMA_B200_STAGE_ADD: launch_add(it); break;
            case DCMA_B200_STAGE_COMPOSE: launch_compose(it); break;
            case DCMA_B200_STAGE_WARP:
                launch_warp_to_buffer();
                w_src_ = 2;
                break;
            default: throw std::invalid_argument("unknown stage id");
        }
        DCMA_CUDA(cudaStreamSynchronize(stream_));
    }

    void get(int buffer, void *host) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        const long long N = g_.N;
        switch (buffer) {
            case DCMA_B200_BUF_FIXED:  // the slices this engine holds
                DCMA_CUDA(cudaMemcpyAsync(host, dF_, sizeof(float) * N * g_.CH, cudaMemcpyDeviceToHost, stream_));
                break;
            case DCMA_B200_BUF_MOVING:  // always the whole volume
                DCMA_CUDA(cudaMemcpyAsync(host, dM_, sizeof(float) * Ng_ * g_.CH, cudaMemcpyDeviceToHost, stream_));
                break;
            case DCMA_B200_BUF_WARPED:
                materialise_warped();
                DCMA_CUDA(cudaMemcpyAsync(host, dW_, sizeof(float) * N * g_.CH, cudaMemcpyDeviceToHost, stream_));
                break;
            case DCMA_B200_BUF_GRADIENT: {
                double *tmp = nullptr;
                DCMA_CUDA(cudaMalloc(&tmp, sizeof(double) * 3 * N));
                k_gradient<<<grid3_, dim3(kTileX, kTileY), 0, stream_>>>(dF_, tmp, g_, zlen_);
                DCMA_CUDA(cudaGetLastError());
                DCMA_CUDA(cudaMemcpyAsync(host, tmp, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, stream_));
                DCMA_CUDA(cudaStreamSynchronize(stream_));
                cudaFree(tmp);
                break;
            }
            case DCMA_B200_BUF_UPDATE:
            case DCMA_B200_BUF_DEFORMATION:
                if (buffer == DCMA_B200_BUF_UPDATE)
                    export_field_host(dU_, static_cast<double *>(host));
                else
                    export_field_host(dD_, static_cast<double *>(host));
                break;
            default: throw std::invalid_argument("unknown buffer id");
        }
        DCMA_CUDA(cudaStreamSynchronize(stream_));
    }

    void set(int buffer, const void *host) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        const long long N = g_.N;
        switch (buffer) {
            case DCMA_B200_BUF_FIXED:
                DCMA_CUDA(cudaMemcpyAsync(dF_, host, sizeof(float) * N * g_.CH, cudaMemcpyHostToDevice, stream_));
                break;
            case DCMA_B200_BUF_MOVING:
                DCMA_CUDA(cudaMemcpyAsync(dM_, host, sizeof(float) * Ng_ * g_.CH, cudaMemcpyHostToDevice, stream_));
                break;
            case DCMA_B200_BUF_WARPED:
                ensure_warped_buffer();
                DCMA_CUDA(cudaMemcpyAsync(dW_, host, sizeof(float) * N * g_.CH, cudaMemcpyHostToDevice, stream_));
                w_src_ = 2;
                break;
            case DCMA_B200_BUF_UPDATE:
            case DCMA_B200_BUF_DEFORMATION:
                if (buffer == DCMA_B200_BUF_UPDATE)
                    import_field_host(static_cast<const double *>(host), dU_);
                else
                    import_field_host(static_cast<const double *>(host), dD_);
                break;
            default: throw std::invalid_argument("buffer is not writable");
        }
        DCMA_CUDA(cudaStreamSynchronize(stream_));
    }

    void export_deformation_device(double *device_out) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        k_export_aos<T><<<grid_lin_, 256, 0, stream_>>>(dD_, device_out, g_.N, owned_v0(), owned_v1());
        DCMA_CUDA(cudaGetLastError());
    }

    // warm start (pyramid): D := the given field; the next force evaluation warps M through it
    void import_deformation_device(const double *device_aos) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        k_import_aos<T><<<grid_lin_, 256, 0, stream_>>>(device_aos, dD_, g_.N, owned_v0(), owned_v1());
        DCMA_CUDA(cudaGetLastError());
        w_src_ = 1;
    }
    void import_deformation_whole(const double *device_aos_whole) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        const long long first = static_cast<long long>(g_.zoff) * g_.R * g_.C;  // global index of local voxel 0
        k_import_aos<T><<<grid_lin_, 256, 0, stream_>>>(device_aos_whole + 3 * first, dD_, g_.N, 0, g_.N);
        DCMA_CUDA(cudaGetLastError());
        w_src_ = 1;
    }
    const float *fixed_dev() const override { return dF_; }
    const float *moving_dev() const override { return dM_; }

    // export_deformation_volume (cc:110-113): SoA T -> the reference's AoS double layout, converted chunk by chunk
    // through two small staging buffers so that no field-sized allocation is made and the conversion of chunk i+1
    // overlaps the PCIe copy of chunk i (the copy engine and the SMs work on different staging halves).
    static constexpr long long kStageVoxels = 1LL << 21;  // 2 Mi voxels = 48 MiB of AoS doubles per half
    double *staging(int half) {
        if (!d_stage_[half]) d_stage_[half] = alloc<double>(3 * kStageVoxels);
        return d_stage_[half];
    }
    void get_deformation_chunked(double *host, void (*on_chunk)(void *, int64_t, int64_t), void *user) override {
        DCMA_CUDA(cudaSetDevice(dev_));
        export_field_host(dD_, host, on_chunk, user);
    }
    template <typename X>
    void export_field_host(const X *field, double *host, void (*on_chunk)(void *, int64_t, int64_t) = nullptr, void *user = nullptr) {
        if (!ev_stage_[0]) {
            DCMA_CUDA(cudaEventCreateWithFlags(&ev_stage_[0], cudaEventDisableTiming));
            DCMA_CUDA(cudaEventCreateWithFlags(&ev_stage_[1], cudaEventDisableTiming));
            DCMA_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
        }
        cudaEvent_t done;
        DCMA_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        int k = 0;
        long long last_v0 = 0, last_v1 = 0;
        const long long vb = owned_v0(), ve = owned_v1();  // owned slices only; the AoS output starts at the first of them
        for (long long v0 = vb; v0 < ve; v0 += kStageVoxels, ++k) {
            const long long v1 = std::min(ve, v0 + kStageVoxels);
            const int half = k & 1;
            double *st = staging(half);
            // the kernel may overwrite this half only after the copy that last read it has finished
            DCMA_CUDA(cudaStreamWaitEvent(stream_, ev_stage_[half], 0));
            k_export_aos<X><<<grid_lin_, 256, 0, stream_>>>(field, st, g_.N, v0, v1);
            DCMA_CUDA(cudaGetLastError());
            DCMA_CUDA(cudaEventRecord(done, stream_));
            DCMA_CUDA(cudaStreamWaitEvent(copy_stream_, done, 0));
            DCMA_CUDA(cudaMemcpyAsync(host + 3 * (v0 - vb), st, sizeof(double) * 3 * (v1 - v0), cudaMemcpyDeviceToHost, copy_stream_));
            DCMA_CUDA(cudaEventRecord(ev_stage_[half], copy_stream_));
            if (on_chunk && k >= 1) {  // the previous chunk's copy: its event is the other half's, recorded one step ago
                DCMA_CUDA(cudaEventSynchronize(ev_stage_[half ^ 1]));
                on_chunk(user, (v0 - kStageVoxels) - vb, v0 - vb);
            }
            last_v0 = v0;
            last_v1 = v1;
        }
        DCMA_CUDA(cudaStreamSynchronize(copy_stream_));
        DCMA_CUDA(cudaStreamSynchronize(stream_));
        if (on_chunk && last_v1 > last_v0) on_chunk(user, last_v0 - vb, last_v1 - vb);
        cudaEventDestroy(done);
    }
    template <typename X>
    void import_field_host(const double *host, X *field) {
        const long long vb = owned_v0(), ve = owned_v1();
        for (long long v0 = vb; v0 < ve; v0 += kStageVoxels) {
            const long long v1 = std::min(ve, v0 + kStageVoxels);
            double *st = staging(0);
            DCMA_CUDA(cudaMemcpyAsync(st, host + 3 * (v0 - vb), sizeof(double) * 3 * (v1 - v0), cudaMemcpyHostToDevice, stream_));
            k_import_aos<X><<<grid_lin_, 256, 0, stream_>>>(st, field, g_.N, v0, v1);
            DCMA_CUDA(cudaGetLastError());
        }
        DCMA_CUDA(cudaStreamSynchronize(stream_));
    }

  private:
    struct ProfRec {
        int stage;
        int launches;
        cudaEvent_t a, b;
    };

    template <typename X>
    X *alloc(long long n) {
        void *p = nullptr;
        const size_t bytes = sizeof(X) * static_cast<size_t>(std::max<long long>(n, 1));
        cudaError_t e = DevicePool::instance().acquire(dev_, bytes, &p);  // device_pool.h: reuses the last engine's blocks
        if (e != cudaSuccess) {
            for (auto &q : allocs_) DevicePool::instance().release(dev_, q.first, q.second);
            allocs_.clear();
            throw CudaError(e, std::string("cudaMalloc of ") + std::to_string(bytes) + " bytes failed: " +
                                   cudaGetErrorString(e));
        }
        allocs_.emplace_back(p, bytes);
        bytes_ += static_cast<int64_t>(bytes);
        return static_cast<X *>(p);
    }

    void reset_ctrl() {
        Ctrl c;
        c.prev_mse = std::numeric_limits<double>::infinity();
        c.last_mse = 0.0;
        c.last_n = 0;
        c.iters_done = 0;
        c.converged_at = kNotConverged;
        c.threshold = p_.convergence_threshold;
        c.ticket = 0;
        c.halo_overflow = 0;
        c.last_sum = 0.0;
        *h_ctrl_ = c;
        DCMA_CUDA(cudaMemcpyAsync(d_ctrl_, h_ctrl_, sizeof(Ctrl), cudaMemcpyHostToDevice, stream_));
        DCMA_CUDA(cudaStreamSynchronize(stream_));  // h_ctrl_ is reused as the D2H landing buffer
    }

    // Gaussian weights and per-position weight sums, cc:574-600 (host side, identical arithmetic and order).
    template <typename X>
    void build_plan(SmoothPlan<X> &pl, double sigma_mm) {
        pl.active = (sigma_mm > 0.0);
        if (!pl.active) return;
        const double pxl[3] = {g_.px, g_.py, g_.pz};
        const int extent[3] = {g_.C, g_.R, g_.Sg};  // the z tables cover the WHOLE volume; kernels get them shifted by zoff
        for (int a = 0; a < 3; ++a) {
            const long long rad = std::max<long long>(1, static_cast<long long>(3.0 * sigma_mm / pxl[a]));
            if (rad > 100000) throw std::invalid_argument("smoothing radius too large");
            pl.r[a] = static_cast<int>(rad);
            const double sigma_pix = sigma_mm / pxl[a];
            std::vector<double> k(2 * rad + 1);
            double sum = 0.0;
            for (long long i = -rad; i <= rad; ++i) {
                const double v = std::exp(-0.5 * (static_cast<double>(i * i)) / (sigma_pix * sigma_pix));
                k[static_cast<size_t>(i + rad)] = v;
                sum += v;
            }
            for (auto &v : k) v /= sum;
            pl.w[a] = k;
            pl.dw[a] = alloc<double>(2 * rad + 1);
            DCMA_CUDA(cudaMemcpy(pl.dw[a], k.data(), sizeof(double) * k.size(), cudaMemcpyHostToDevice));
        }
        // The fused kernel serves any per-axis radii up to kMaxFastRadius (anisotropic clinical spacing included): it
        // is instantiated for rk = the largest of the three and the other axes' weights are zero-padded.
        pl.rk = std::max(pl.r[0], std::max(pl.r[1], pl.r[2]));
        pl.fast = (p_.fused != 0) && (pl.rk <= kMaxFastRadius);
        if (pl.fast) {
            const int rk = pl.rk;
            for (int a = 0; a < 3; ++a) {
                const int rad = pl.r[a];
                pl.W.r[a] = rad;
                for (int k = 0; k < 2 * kMaxFastRadius + 1; ++k) pl.W.w[a][k] = X(0);
                for (int k = -rad; k <= rad; ++k) pl.W.w[a][k + rk] = static_cast<X>(pl.w[a][k + rad]);
                const int padded = ((extent[a] + 127) / 128) * 128 + 128;  // whole tiles: the kernel never tests bounds
                std::vector<X> tab(padded, X(1));
                for (int pos = 0; pos < extent[a]; ++pos) {
                    // in-order accumulation over the in-grid taps, in the kernel's own type (cc:620-631); the padded
                    // zero weights leave the sum unchanged
                    X wsum = 0;
                    for (int k = -rk; k <= rk; ++k) {
                        const int q = pos + k;
                        if (q >= 0 && q < extent[a]) wsum += pl.W.w[a][k + rk];
                    }
                    tab[pos] = STRICT ? wsum : static_cast<X>(1.0 / static_cast<double>(wsum));
                }
                pl.tab[a] = alloc<X>(padded);
                DCMA_CUDA(cudaMemcpy(pl.tab[a], tab.data(), sizeof(X) * tab.size(), cudaMemcpyHostToDevice));
            }
        }
    }

    // plans of the stage API, per element type of the field being smoothed
    template <typename X>
    SmoothPlan<X> &plan_for(double sigma_mm) {
        if constexpr (std::is_same<X, TU>::value) {
            if (sigma_mm == p_.update_field_smoothing_sigma) return plan_u_;
        }
        if constexpr (std::is_same<X, T>::value) {
            if (sigma_mm == p_.deformation_field_smoothing_sigma) return plan_d_;
        }
        auto &plans = extra_plans<X>();
        auto it = plans.find(sigma_mm);
        if (it == plans.end()) {
            it = plans.emplace(sigma_mm, SmoothPlan<X>()).first;
            build_plan(it->second, sigma_mm);
        }
        return it->second;
    }
    template <typename X>
    std::map<double, SmoothPlan<X>> &extra_plans() {
        if constexpr (std::is_same<X, T>::value)
            return extra_plans_;
        else
            return extra_plans_u_;
    }
    // scratch field of the update-field smoothing: D's scratch when both fields have one type
    TU *&tmp_u() {
        if constexpr (kSameU)
            return dTmp_;
        else
            return dTmpU_;
    }

    void prof_begin(int stage) {
        if (!profiling_) return;
        ProfRec r;
        r.stage = stage;
        r.launches = 0;
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, stream_);
        prof_.push_back(r);
        prof_launch_mark_ = launches_;
    }
    void prof_end() {
        if (!profiling_) return;
        cudaEventRecord(prof_.back().b, stream_);
        prof_.back().launches = static_cast<int>(launches_ - prof_launch_mark_);
    }

    // compute_single_iteration, cc:81-108, with warp_moving moved to the front of the NEXT iteration's force kernel.
    void iteration(long long it, double *d_hist) {
        const bool diffeo = (p_.use_diffeomorphic != 0);
        if (p_.use_symmetric_force) {
            // the symmetric force needs the gradient of W, i.e. W in memory with its neighbours (extension, row a11)
            prof_begin(DCMA_B200_STAGE_WARP);
            materialise_warped();
            prof_end();
            prof_begin(DCMA_B200_STAGE_FORCE);
            launch_force(2, it, d_hist, it, 1, /*reduce=*/false, /*fused_reduce=*/true);
            prof_end();
        } else if (p_.fused) {
            prof_begin(DCMA_B200_STAGE_FORCE);
            launch_force(w_src_, it, d_hist, it, 1, /*reduce=*/false, /*fused_reduce=*/DCMA_FUSED_REDUCE != 0);
            prof_end();
            if (!DCMA_FUSED_REDUCE) {
                prof_begin(DCMA_B200_STAGE_REDUCE);
                launch_reduce(it, d_hist, it, 1);
                prof_end();
            }
        } else {
            if (w_src_ == 1) {
                prof_begin(DCMA_B200_STAGE_WARP);
                launch_warp_to_buffer(it);
                prof_end();
                w_src_ = 2;
            }
            prof_begin(DCMA_B200_STAGE_FORCE);
            launch_force(w_src_, it, d_hist, it, 1, /*reduce=*/false);
            prof_end();
        }
        if (!p_.fused && !p_.use_symmetric_force) {
            prof_begin(DCMA_B200_STAGE_REDUCE);
            launch_reduce(it, d_hist, it, 1);
            prof_end();
        }
        if (diffeo && plan_u_.active) {
            prof_begin(DCMA_B200_STAGE_SMOOTH_U);
            smooth(dU_, tmp_u(), plan_u_, it);
            prof_end();
        }
        if (!diffeo) {
            prof_begin(DCMA_B200_STAGE_ADD);
            launch_add(it);
            prof_end();
        } else {
            prof_begin(DCMA_B200_STAGE_COMPOSE);
            launch_compose(it);
            prof_end();
        }
        if (plan_d_.active) {
            prof_begin(DCMA_B200_STAGE_SMOOTH_D);
            smooth(dD_, dTmp_, plan_d_, it);
            prof_end();
        }
        w_src_ = 1;  // the next force kernel warps M through the new D (cc:106)
        record_rotation(it);
    }

    // which allocation plays D / U / scratch after loop iteration `it` (the smoothing calls swap pointers)
    struct Rotation {
        T *d;
        TU *u;
        T *tmp;
        TU *tmpu;
    };
    void record_rotation(long long it) {
        if (it < 0) return;
        if (rot_hist_.size() <= static_cast<size_t>(it)) rot_hist_.resize(static_cast<size_t>(it) + 1, Rotation{dD_, dU_, dTmp_, dTmpU_});
        rot_hist_[static_cast<size_t>(it)] = Rotation{dD_, dU_, dTmp_, dTmpU_};
    }

    // zr0/zr1 >= 0: only the owned slices [zr0, zr1) (local indices), MSE partials written from `part_off` on (the caller
    // reduces all parts with launch_reduce(..., nparts)); returns the number of partials this launch wrote
    int launch_force(int w_src, long long it, double *d_hist, long long hist_idx, int count_iter, bool reduce = true,
                     bool fused_reduce = false, int zr0 = -1, int zr1 = -1, int part_off = 0) {
        if (w_src == 2) ensure_warped_buffer();
        ForceK fk;
        fk.norm_eps = p_.normalization_factor + 1.0e-10;  // cc:315
        fk.inv_norm_eps = 1.0 / fk.norm_eps;
        fk.max_update = p_.max_update_magnitude;
        fk.mu2_lo = (fk.max_update >= 0.0) ? fk.max_update * fk.max_update * (1.0 - 8.881784197001252e-16) : -1.0;
        const dim3 blk(kTileX, kTileY);
        Geo gq = g_;
        dim3 grid = grid3_;
        if (zr0 >= 0) {
            gq.zb = zr0;
            gq.ze = zr1;
            grid.z = static_cast<unsigned>((zr1 - zr0 + zlen_ - 1) / zlen_);
            if (zr1 <= zr0) return 0;
        }
        const int nparts = static_cast<int>(grid.x * grid.y * grid.z);
        if (part_off + nparts > part_cap_) throw std::logic_error("force stage split into too many parts");
        double *psum = d_part_sum_ + part_off;
        long long *pcnt = d_part_cnt_ + part_off;
#define DCMA_LAUNCH_FORCE2(IX, CH1, SYM)                                                                             \
    k_warp_force<T, STRICT, IX, CH1, SYM, TU><<<grid, blk, 0, stream_>>>(dF_, dM_, dW_, dD_, dU_, gq, zlen_, w_src, fk, psum, \
                                                                       pcnt, d_ctrl_, it, fused_reduce ? 1 : 0, d_hist, \
                                                                       hist_idx, count_iter)
#define DCMA_LAUNCH_FORCE(IX, CH1)                            \
    do {                                                      \
        if (p_.use_symmetric_force && w_src == 2)             \
            DCMA_LAUNCH_FORCE2(IX, CH1, true);                \
        else                                                  \
            DCMA_LAUNCH_FORCE2(IX, CH1, false);               \
    } while (0)
        if (small_index_) {
            if (g_.CH == 1) DCMA_LAUNCH_FORCE(int, true); else DCMA_LAUNCH_FORCE(int, false);
        } else {
            if (g_.CH == 1) DCMA_LAUNCH_FORCE(long long, true); else DCMA_LAUNCH_FORCE(long long, false);
        }
#undef DCMA_LAUNCH_FORCE2
#undef DCMA_LAUNCH_FORCE
        DCMA_CUDA(cudaGetLastError());
        ++launches_;
        if (reduce) launch_reduce(it, d_hist, hist_idx, count_iter && d_hist != nullptr ? 1 : 0);
        return nparts;
    }
    void launch_reduce(long long it, double *d_hist, long long hist_idx, int count_iter, int nparts = -1) {
        k_reduce_final<<<1, 256, 0, stream_>>>(d_part_sum_, d_part_cnt_, nparts < 0 ? grid_vox_ : nparts, d_ctrl_, d_hist, it,
                                               hist_idx, count_iter);
        DCMA_CUDA(cudaGetLastError());
        ++launches_;
    }
    void launch_add(long long it) {
        k_add_update<T, TU><<<grid_lin_, 256, 0, stream_>>>(dD_, dU_, g_.N, g_.px, g_.py, g_.pz, d_ctrl_, it);
        DCMA_CUDA(cudaGetLastError());
        ++launches_;
    }
    void launch_compose(long long it, int zr0 = -1, int zr1 = -1) {
        Geo gq = g_;
        dim3 grid = grid3_;
        if (zr0 >= 0) {
            if (zr1 <= zr0) return;
            gq.zb = zr0;
            gq.ze = zr1;
            grid.z = static_cast<unsigned>((zr1 - zr0 + zlen_ - 1) / zlen_);
        }
        if (small_index_)
            k_compose<T, STRICT, int, TU><<<grid, dim3(kTileX, kTileY), 0, stream_>>>(dD_, dU_, gq, zlen_, d_ctrl_, it);
        else
            k_compose<T, STRICT, long long, TU><<<grid, dim3(kTileX, kTileY), 0, stream_>>>(dD_, dU_, gq, zlen_, d_ctrl_, it);
        DCMA_CUDA(cudaGetLastError());
        ++launches_;
    }
    void ensure_warped_buffer() {
        if (!dW_) {
            dW_ = alloc<float>(g_.N * g_.CH);
            DCMA_CUDA(cudaMemcpyAsync(dW_, dM_ + local_offset_in_moving(), sizeof(float) * g_.N * g_.CH, cudaMemcpyDeviceToDevice, stream_));
        }
    }
    // z_margin: also warp that many held planes on either side of the owned range (slab engines, symmetric force)
    void launch_warp_to_buffer(long long it = 0, int z_margin = 0) {
        (void)it;
        ensure_warped_buffer();
        const dim3 blk(kTileX, kTileY);
        Geo gq = g_;
        dim3 grid = grid3_;
        if (z_margin > 0) {
            gq.zb = std::max(0, g_.zb - z_margin);
            gq.ze = std::min(g_.S, g_.ze + z_margin);
            grid.z = static_cast<unsigned>((gq.ze - gq.zb + zlen_ - 1) / zlen_);
        }
        if (small_index_) {
            if (g_.CH == 1)
                k_warp<T, STRICT, int, true><<<grid, blk, 0, stream_>>>(dM_, dD_, dW_, gq, zlen_, 0, 0.f);
            else
                k_warp<T, STRICT, int, false><<<grid, blk, 0, stream_>>>(dM_, dD_, dW_, gq, zlen_, 0, 0.f);
        } else {
            if (g_.CH == 1)
                k_warp<T, STRICT, long long, true><<<grid, blk, 0, stream_>>>(dM_, dD_, dW_, gq, zlen_, 0, 0.f);
            else
                k_warp<T, STRICT, long long, false><<<grid, blk, 0, stream_>>>(dM_, dD_, dW_, gq, zlen_, 0, 0.f);
        }
        DCMA_CUDA(cudaGetLastError());
        ++launches_;
    }
    // bring the explicit `warped` buffer up to date with the engine state
    void materialise_warped(int z_margin = 0) {
        ensure_warped_buffer();
        if (w_src_ == 0) {
            DCMA_CUDA(cudaMemcpyAsync(dW_, dM_ + local_offset_in_moving(), sizeof(float) * g_.N * g_.CH, cudaMemcpyDeviceToDevice, stream_));
        } else if (w_src_ == 1) {
            launch_warp_to_buffer(0, z_margin);
        }
        w_src_ = 2;
    }

    // smooth_on_device, cc:563-702.  Result ends in `field` (pointer swap instead of the reference's copy-back).
    template <typename X>
    void smooth(X *&field, X *&tmp, SmoothPlan<X> &pl, long long it) {
        if (!pl.active) return;
        if (pl.fast) {
            launch_smooth3d(field, tmp, pl, it);
            std::swap(field, tmp);
        } else {
            X *bufs[2] = {field, tmp};
            for (int a = 0; a < 3; ++a) {  // x: field->tmp (cc:609), y: tmp->field (cc:638), z: field->tmp (cc:667)
                k_smooth_axis<X><<<grid_lin_, 256, 0, stream_>>>(bufs[a & 1], bufs[(a + 1) & 1], g_, a, pl.r[a], pl.dw[a],
                                                              d_ctrl_, it);
                DCMA_CUDA(cudaGetLastError());
                ++launches_;
            }
            // three passes: field -> tmp -> field -> tmp; the result is in tmp
            std::swap(field, tmp);
        }
    }

    template <typename X, int R, int TX, int TY, int NT, int MINB, bool VEC>
    void launch_smooth3d_cfg(const X *in, X *out, const SmoothPlan<X> &pl, long long it) {
        using Cfg = Smooth3dCfg<X, R, TX, TY, NT>;
        constexpr bool WS = (DCMA_SMOOTH_WS < 0) ? (sizeof(X) == 8) : (DCMA_SMOOTH_WS != 0);
        constexpr int NTL = WS ? 2 * NT : NT;  // threads per CTA at launch
        constexpr size_t SMEM = WS ? Cfg::smem_bytes_ws : Cfg::smem_bytes;
        auto kern = k_smooth3d<X, STRICT, R, TX, TY, NT, (WS ? 2 : MINB), VEC, WS>;
        static thread_local bool attr_set[64] = {false};
        if (!attr_set[dev_ & 63]) {
            DCMA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SMEM)));
            // without this the driver picks the smallest L1/shared split that fits ONE block (ncu: 1 block/SM)
            DCMA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared));
            attr_set[dev_ & 63] = true;
        }
        const int gx = (g_.C + TX - 1) / TX, gy = (g_.R + TY - 1) / TY;
        // z chunks: score = (fraction of slice steps that produce output) x (wave efficiency); every chunk re-stages
        // 2R slices of warm-up, and a partially filled last wave idles SMs.
        int blocks_per_sm = 1;
        DCMA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, NTL, SMEM));
        const double slots = static_cast<double>(std::max(1, blocks_per_sm)) * sms_;
        const long long nxy = static_cast<long long>(gx) * gy * 3;
        int best_nz = 1;
        double best_score = -1.0;
        const int So = g_.ze - g_.zb;  // owned slices
        const int nz_max = std::max(1, std::min(So / std::max(2 * R, 1), static_cast<int>(65535 / 3)));
        for (int nzc = 1; nzc <= nz_max; ++nzc) {
            const int zc = (So + nzc - 1) / nzc;
            const int nz_eff = (So + zc - 1) / zc;
            const double waves = static_cast<double>(nxy) * nz_eff / slots;
            const double score = (static_cast<double>(zc) / (zc + 2 * R)) * (waves / std::ceil(waves));
            if (score > best_score + 1e-9) {
                best_score = score;
                best_nz = nz_eff;
            }
        }
        const int zchunk = (So + best_nz - 1) / best_nz;
        const int nz = (So + zchunk - 1) / zchunk;
        if (3LL * nz > 65535 || gy > 65535) throw std::invalid_argument("volume exceeds the smoothing kernel's grid limits");
        kern<<<dim3(gx, gy, 3 * nz), NTL, SMEM, stream_>>>(in, out, g_, pl.W, pl.tab[0], pl.tab[1], pl.tab[2] + g_.zoff,
                                                                     pl.dw[0], pl.dw[1], pl.dw[2], zchunk, nz, d_ctrl_, it);
        DCMA_CUDA(cudaGetLastError());
        ++launches_;
    }


    // ---- k_smooth3d_tma (fp64 fields): TMA descriptor per (field allocation, box shape) and the launch ---------------
    using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                       const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn encode_tiled_fn() {
        static EncodeTiledFn fn = []() -> EncodeTiledFn {
            void *p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
                q != cudaDriverEntryPointSuccess)
                return nullptr;
            return reinterpret_cast<EncodeTiledFn>(p);
        }();
        return fn;
    }
    // the field as a 4-D tensor {x, y, z, component}; one box = one staged slice tile (+halo), out-of-grid -> zeros
    template <typename X>
    const CUtensorMap *tensor_map(const X *base, int box_x, int box_y) {
        const auto key = std::make_tuple(static_cast<const void *>(base), box_x, box_y);
        auto itm = tmaps_.find(key);
        if (itm != tmaps_.end()) return &itm->second;
        EncodeTiledFn enc = encode_tiled_fn();
        if (!enc) return nullptr;
        CUtensorMap m;
        const cuuint64_t dims[4] = {static_cast<cuuint64_t>(g_.C), static_cast<cuuint64_t>(g_.R), static_cast<cuuint64_t>(g_.S), 3};
        const cuuint64_t strides[3] = {static_cast<cuuint64_t>(g_.C) * sizeof(X),
                                       static_cast<cuuint64_t>(g_.R) * g_.C * sizeof(X),
                                       static_cast<cuuint64_t>(g_.N) * sizeof(X)};
        const cuuint32_t box[4] = {static_cast<cuuint32_t>(box_x), static_cast<cuuint32_t>(box_y), 1, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<X *>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return nullptr;
        return &tmaps_.emplace(key, m).first->second;
    }
    template <typename X>
    bool tma_usable(const X *in) const {
        return DCMA_SMOOTH_TMA != 0 && sizeof(X) == 8 && (g_.C % 2) == 0 && (reinterpret_cast<uintptr_t>(in) % 16) == 0 &&
               (g_.N % 2) == 0 && g_.C >= 2;
    }
    template <typename X, int R>
    bool launch_smooth3d_tma(const X *in, X *out, const SmoothPlan<X> &pl, long long it) {
        if constexpr (sizeof(X) != 8 || R > 4) {
            return false;
        } else {
            constexpr int TX = DCMA_TMA_TX, TY = DCMA_TMA_TY, OY = DCMA_TMA_OY, NS = DCMA_TMA_NS, NXB = DCMA_TMA_NXB;
            using Cfg = SmoothTmaCfg<X, R, TX, TY, OY, NS, NXB>;
            if (!tma_usable(in)) return false;
            const CUtensorMap *tm = tensor_map(in, Cfg::PWP, Cfg::PH);
            if (!tm) return false;
            auto kern = k_smooth3d_tma<X, STRICT, R, TX, TY, OY, NS, NXB>;
            static thread_local bool attr_set[64] = {false};
            if (!attr_set[dev_ & 63]) {
                DCMA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Cfg::smem_bytes)));
                DCMA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                attr_set[dev_ & 63] = true;
            }
            const int gx = (g_.C + TX - 1) / TX, gy = (g_.R + TY - 1) / TY;
            int blocks_per_sm = 1;
            DCMA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, Cfg::NTHREADS, Cfg::smem_bytes));
            // persistent launch, statically balanced (see SmoothWork): every CTA gets the same number of slice steps
            const long long ncols = 3LL * gx * gy;
            if (ncols > 0x7fffffffLL) return false;
            const int So = g_.ze - g_.zb;
            const int G = static_cast<int>(std::min<long long>(ncols, static_cast<long long>(std::max(1, blocks_per_sm)) * sms_));
            SmoothWork work;
            work.ncols = static_cast<int>(ncols);
            work.tiles_x = gx;
            work.tiles_y = gy;
            work.nfull = static_cast<int>(ncols / G);
            const long long rem = (ncols - static_cast<long long>(work.nfull) * G) * So;
            work.per = static_cast<int>((rem + G - 1) / G);
            // pieces shorter than the warm-up are not worth a CTA of their own: fall back to whole columns
            if (work.per > 0 && work.per < 2 * R) {
                work.per = So;
            }
            kern<<<G, Cfg::NTHREADS, Cfg::smem_bytes, stream_>>>(*tm, in, out, g_, pl.W, pl.tab[0], pl.tab[1], pl.tab[2] + g_.zoff,
                                                                 pl.dw[0], pl.dw[1], pl.dw[2], work, d_ctrl_, it);
            DCMA_CUDA(cudaGetLastError());
            ++launches_;
            return true;
        }
    }
    // z chunks: score = (fraction of slice steps that produce output) x (wave efficiency); every chunk re-stages
    // 2R slices of warm-up, and a partially filled last wave idles SMs.
    void choose_zchunks(int gx, int gy, int R, int blocks_per_sm, int &zchunk, int &nz) const {
        const double slots = static_cast<double>(std::max(1, blocks_per_sm)) * sms_;
        const long long nxy = static_cast<long long>(gx) * gy * 3;
        int best_nz = 1;
        double best_score = -1.0;
        const int So = g_.ze - g_.zb;  // owned slices
        const int nz_max = std::max(1, std::min(So / std::max(2 * R, 1), static_cast<int>(65535 / 3)));
        for (int nzc = 1; nzc <= nz_max; ++nzc) {
            const int zc = (So + nzc - 1) / nzc;
            const int nz_eff = (So + zc - 1) / zc;
            const double waves = static_cast<double>(nxy) * nz_eff / slots;
            const double score = (static_cast<double>(zc) / (zc + 2 * R)) * (waves / std::ceil(waves));
            if (score > best_score + 1e-9) {
                best_score = score;
                best_nz = nz_eff;
            }
        }
        zchunk = (So + best_nz - 1) / best_nz;
        nz = (So + zchunk - 1) / zchunk;
    }

    template <typename X, int R>
    void launch_smooth3d_r(const X *in, X *out, const SmoothPlan<X> &pl, long long it) {
        // NT = 128 threads per role, 2 adjacent output rows x one 16-byte vector per thread; the register z window
        // (72 registers) puts the one-role kernel at ~165 registers, so 3 CTAs/SM (fp32); the warp-specialised form
        // (fp64) launches 2*NT threads, 2 CTAs/SM.  float: 64x16 tile; double: 32x16 tile.
        constexpr int TX = (sizeof(X) == 4) ? 64 : 32;
        constexpr int TY = DCMA_SMOOTH_TY;
        constexpr int NT = DCMA_SMOOTH_NT;
        constexpr int MINB = DCMA_SMOOTH_MINB;
        if (launch_smooth3d_tma<X, R>(in, out, pl, it)) return;
        const bool vec = (g_.C % static_cast<int>(16 / sizeof(X)) == 0);
        if (vec)
            launch_smooth3d_cfg<X, R, TX, TY, NT, MINB, true>(in, out, pl, it);
        else
            launch_smooth3d_cfg<X, R, TX, TY, NT, MINB, false>(in, out, pl, it);
    }

    template <typename X>
    void launch_smooth3d(const X *in, X *out, const SmoothPlan<X> &pl, long long it) {
        switch (pl.rk) {
            case 1: launch_smooth3d_r<X, 1>(in, out, pl, it); break;
            case 2: launch_smooth3d_r<X, 2>(in, out, pl, it); break;
            case 3: launch_smooth3d_r<X, 3>(in, out, pl, it); break;
            case 4: launch_smooth3d_r<X, 4>(in, out, pl, it); break;
            case 5: launch_smooth3d_r<X, 5>(in, out, pl, it); break;
            case 6: launch_smooth3d_r<X, 6>(in, out, pl, it); break;
            default: throw std::logic_error("fast smoothing radius out of range");
        }
    }

    bool is_slab() const { return g_.S != g_.Sg || g_.zb != 0 || g_.ze != g_.S; }
    long long owned_v0() const { return static_cast<long long>(g_.zb) * g_.R * g_.C; }
    long long owned_v1() const { return static_cast<long long>(g_.ze) * g_.R * g_.C; }
    long long local_offset_in_moving() const { return static_cast<long long>(g_.zoff) * g_.R * g_.C * g_.CH; }

    Geo g_;
    long long Ng_ = 0;  // voxels of the whole volume
    SlabSpec slab_;
    double *d_hist_ = nullptr;
    size_t d_hist_bytes_ = 0;
    void drop_history() noexcept {
        if (d_hist_) DevicePool::instance().release(dev_, d_hist_, d_hist_bytes_);
        d_hist_ = nullptr;
    }
    cudaEvent_t ev_run0_ = nullptr, ev_run1_ = nullptr;
    dcma_b200_demons_params p_;
    int dev_ = 0, sms_ = 148;
    cudaStream_t stream_ = nullptr;
    bool own_stream_ = false;
    bool loaded_ = false;
    dim3 grid3_;
    bool small_index_ = true;
    int zlen_ = 1;
    int grid_vox_ = 1, grid_lin_ = 1, part_cap_ = 1;
    static constexpr int kMaxForceParts = 4;
    int force_parts_ = 0;  // partials written by the force_range calls of the current iteration
    float *dF_ = nullptr, *dM_ = nullptr, *dW_ = nullptr;
    T *dD_ = nullptr, *dTmp_ = nullptr;
    TU *dU_ = nullptr, *dTmpU_ = nullptr;  // dTmpU_: only when TU != T
    void *field_alloc_[4] = {nullptr, nullptr, nullptr, nullptr};
    int n_field_alloc_ = 3;
    double *d_part_sum_ = nullptr;
    long long *d_part_cnt_ = nullptr;
    Ctrl *d_ctrl_ = nullptr, *h_ctrl_ = nullptr;
    double *d_stage_[2] = {nullptr, nullptr};
    cudaEvent_t ev_stage_[2] = {nullptr, nullptr};
    cudaStream_t copy_stream_ = nullptr;
    SmoothPlan<TU> plan_u_;
    SmoothPlan<T> plan_d_;
    std::map<double, SmoothPlan<T>> extra_plans_;
    std::map<double, SmoothPlan<TU>> extra_plans_u_;  // used only when TU != T
    std::map<std::tuple<const void *, int, int>, CUtensorMap> tmaps_;
    int w_src_ = 0;
    std::vector<std::pair<void *, size_t>> allocs_;  // (block, bytes): returned to the device pool, not to the driver
    int64_t bytes_ = 0;
    int64_t launches_ = 0, prof_launch_mark_ = 0;
    bool profiling_ = false;
    std::vector<ProfRec> prof_;
    std::vector<Rotation> rot_hist_;
};

}  // namespace DCMA_NS
