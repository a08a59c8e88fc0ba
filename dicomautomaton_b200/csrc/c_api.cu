// dcma_b200 -- the C ABI (include/dcma_b200/demons_c_api.h).  Nothing throws across this boundary.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <limits>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"
#include "device_pool.h"
#include "engine.h"
#include "pyramid.h"
#include "slab.h"

struct dcma_b200_engine {
    dcma::EngineBase *impl;
    dcma_b200_geom geom;
};
struct dcma_b200_slab_group {
    dcma::SlabGroup *impl;
};

namespace dcma {
EngineBase *make_engine(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream, const SlabSpec &slab) {
    if (p.pyramid_levels > 1)
        throw std::invalid_argument("pyramid_levels > 1 is a property of dcma_b200_demons_register, not of one engine");
    switch (p.field_mode) {
        case DCMA_B200_FIELD_FP64: return make_engine_f64(g, p, stream, slab);
        case DCMA_B200_FIELD_FP64_STRICT: return make_engine_f64_strict(g, p, stream, slab);
        case DCMA_B200_FIELD_FP32: return make_engine_f32(g, p, stream, slab);
        case DCMA_B200_FIELD_MIXED: return make_engine_mixed(g, p, stream, slab);
        default: throw std::invalid_argument("unknown field_mode");
    }
}
}  // namespace dcma

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg, char *errbuf, size_t errbuf_len) {
    g_last_error = msg;
    if (errbuf && errbuf_len > 0) {
        std::snprintf(errbuf, errbuf_len, "%s", msg.c_str());
    }
    return code;
}

template <typename F>
int guarded(F &&f, char *errbuf = nullptr, size_t errbuf_len = 0) {
    try {
        f();
        return DCMA_B200_OK;
    } catch (const dcma::CudaError &e) {
        const int code = (e.code == cudaErrorNoDevice || e.code == cudaErrorInvalidDevice ||
                          e.code == cudaErrorInsufficientDriver)
                             ? DCMA_B200_ENODEVICE
                             : (e.code == cudaErrorMemoryAllocation ? DCMA_B200_ENOMEM : DCMA_B200_ECUDA);
        return fail(code, e.what(), errbuf, errbuf_len);
    } catch (const std::invalid_argument &e) {
        return fail(DCMA_B200_EINVAL, e.what(), errbuf, errbuf_len);
    } catch (const std::logic_error &e) {
        return fail(DCMA_B200_ESTATE, e.what(), errbuf, errbuf_len);
    } catch (const std::bad_alloc &) {
        return fail(DCMA_B200_ENOMEM, "host allocation failed", errbuf, errbuf_len);
    } catch (const std::exception &e) {
        return fail(DCMA_B200_ECUDA, e.what(), errbuf, errbuf_len);
    } catch (...) {
        return fail(DCMA_B200_ECUDA, "unknown error", errbuf, errbuf_len);
    }
}

dcma::EngineBase *make(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream) {
    return dcma::make_engine(g, p, stream, dcma::SlabSpec{});
}

}  // namespace

extern "C" {

int dcma_b200_abi_version(void) { return DCMA_B200_ABI_VERSION; }

const char *dcma_b200_last_error(void) { return g_last_error.c_str(); }

void dcma_b200_device_pool_trim(void) { dcma::DevicePool::instance().trim(); }

int64_t dcma_b200_device_pool_bytes(void) { return static_cast<int64_t>(dcma::DevicePool::instance().cached_bytes()); }

int dcma_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void dcma_b200_demons_params_default(dcma_b200_demons_params *p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->max_iterations = 100;  // reference src/Alignment_Demons.h:23-60
    p->convergence_threshold = 0.001;
    p->deformation_field_smoothing_sigma = 1.0;
    p->update_field_smoothing_sigma = 0.5;
    p->use_diffeomorphic = 0;
    p->use_histogram_matching = 0;
    p->histogram_bins = 256;
    p->histogram_outlier_fraction = 0.01;
    p->normalization_factor = 1.0;
    p->max_update_magnitude = 2.0;
    p->verbosity = 1;
    p->field_mode = DCMA_B200_FIELD_FP64;
    p->device = -1;
    p->check_every = 0;
    p->pyramid_levels = 1;
    p->use_symmetric_force = 0;
    p->fused = 1;
    p->n_devices = 0;
    p->halo_planes = 0;
    p->pyramid_iterations[0] = p->pyramid_iterations[1] = p->pyramid_iterations[2] = 0;
}

int dcma_b200_engine_create(dcma_b200_engine **out, const dcma_b200_geom *geom, const dcma_b200_demons_params *params,
                            void *stream, char *errbuf, size_t errbuf_len) {
    if (!out || !geom || !params) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    *out = nullptr;
    return guarded(
        [&]() {
            dcma::EngineBase *impl = make(*geom, *params, stream);
            *out = new dcma_b200_engine{impl, *geom};
        },
        errbuf, errbuf_len);
}

void dcma_b200_engine_destroy(dcma_b200_engine *e) {
    if (!e) return;
    try {
        delete e->impl;
    } catch (...) {
    }
    delete e;
}

int dcma_b200_engine_load(dcma_b200_engine *e, const float *fixed, const float *moving, int on_device) {
    if (!e || !fixed || !moving) return fail(DCMA_B200_EINVAL, "null argument", nullptr, 0);
    return guarded([&]() { e->impl->load(fixed, moving, on_device != 0); });
}

int dcma_b200_engine_run(dcma_b200_engine *e, int64_t max_iterations, int profile, double *mse_history,
                         dcma_b200_run_stats *stats) {
    if (!e) return fail(DCMA_B200_EINVAL, "null engine", nullptr, 0);
    return guarded([&]() { e->impl->run(max_iterations, profile != 0, mse_history, stats); });
}

int dcma_b200_engine_stage(dcma_b200_engine *e, int stage, double sigma_mm, double *mse_out, int64_t *n_valid_out) {
    if (!e) return fail(DCMA_B200_EINVAL, "null engine", nullptr, 0);
    return guarded([&]() { e->impl->stage(stage, sigma_mm, mse_out, n_valid_out); });
}

int dcma_b200_engine_get(dcma_b200_engine *e, int buffer, void *host) {
    if (!e || !host) return fail(DCMA_B200_EINVAL, "null argument", nullptr, 0);
    return guarded([&]() { e->impl->get(buffer, host); });
}

int dcma_b200_engine_set(dcma_b200_engine *e, int buffer, const void *host) {
    if (!e || !host) return fail(DCMA_B200_EINVAL, "null argument", nullptr, 0);
    return guarded([&]() { e->impl->set(buffer, host); });
}

int dcma_b200_engine_export_deformation_device(dcma_b200_engine *e, double *device_out) {
    if (!e || !device_out) return fail(DCMA_B200_EINVAL, "null argument", nullptr, 0);
    return guarded([&]() { e->impl->export_deformation_device(device_out); });
}

int dcma_b200_engine_sync(dcma_b200_engine *e) {
    if (!e) return fail(DCMA_B200_EINVAL, "null engine", nullptr, 0);
    return guarded([&]() { e->impl->sync(); });
}

int64_t dcma_b200_engine_device_bytes(const dcma_b200_engine *e) { return e ? e->impl->device_bytes() : 0; }

int dcma_b200_demons_register(const dcma_b200_geom *geom, const float *fixed, const float *moving,
                              const dcma_b200_demons_params *params, double *deformation_out, double *mse_history,
                              dcma_b200_run_stats *stats, char *errbuf, size_t errbuf_len) {
    if (!geom || !fixed || !moving || !params || !deformation_out)
        return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    if (stats) std::memset(stats, 0, sizeof(*stats));
    return guarded(
        [&]() {
            using clk = std::chrono::steady_clock;
            struct Holder {
                dcma::EngineBase *p = nullptr;
                ~Holder() { delete p; }
            } h;
            if (params->pyramid_levels > 1) {
                const auto s0 = clk::now();
                dcma_b200_run_stats local;
                dcma::pyramid_register(*geom, fixed, moving, *params, deformation_out, mse_history, &local);
                if (stats) {
                    *stats = local;
                    const long long n = geom->slices * geom->rows * geom->cols;
                    stats->h2d_bytes = 2 * n * geom->channels * static_cast<long long>(sizeof(float));
                    stats->d2h_bytes = 3 * n * static_cast<long long>(sizeof(double));
                    stats->h2d_ms = std::chrono::duration<double, std::milli>(clk::now() - s0).count() - local.loop_ms;
                }
                return;
            }
            if (params->n_devices > 1) {
                // the volume is sharded into z-slabs over devices 0..n_devices-1 of this process
                const std::vector<int> dev = dcma::slab_devices(params->n_devices);
                // A composition gather that leaves the halo fails the run loudly (std::logic_error from SlabGroup::run).
                // With the default halo the registration is then repeated with a doubled one (an explicit halo_planes
                // is the caller's decision and is not second-guessed).
                int64_t halo = params->halo_planes > 0 ? params->halo_planes : dcma::SlabGroup::default_halo(*geom, *params);
                std::chrono::steady_clock::time_point s0, s1, s2, s3;
                dcma_b200_run_stats local;
                for (;;) {
                    dcma::SlabGroup grp(*geom, *params, dev, halo);
                    s0 = clk::now();
                    grp.load(fixed, moving, false);
                    s1 = clk::now();
                    try {
                        grp.run(params->max_iterations, mse_history, &local);
                    } catch (const std::logic_error &ex) {
                        const bool overflow = std::string(ex.what()).find("left the slab halo") != std::string::npos;
                        if (!overflow || params->halo_planes > 0 || halo * 2 * params->n_devices > geom->slices) throw;
                        if (params->verbosity >= 1)
                            std::fprintf(stderr, "[dcma_b200] %s -- repeating with halo_planes = %lld\n", ex.what(), (long long)(2 * halo));
                        halo *= 2;
                        continue;
                    }
                    s2 = clk::now();
                    grp.get_deformation(deformation_out);
                    s3 = clk::now();
                    break;
                }
                if (stats) {
                    *stats = local;
                    const long long n = geom->slices * geom->rows * geom->cols;
                    stats->h2d_ms = std::chrono::duration<double, std::milli>(s1 - s0).count();
                    stats->d2h_ms = std::chrono::duration<double, std::milli>(s3 - s2).count();
                    stats->h2d_bytes = (n + n * params->n_devices) * geom->channels * static_cast<long long>(sizeof(float));
                    stats->d2h_bytes = 3 * n * static_cast<long long>(sizeof(double));
                }
                return;
            }
            const auto tc = clk::now();
            h.p = make(*geom, *params, nullptr);
            const auto t0 = clk::now();
            h.p->load(fixed, moving, false);
            const auto t1 = clk::now();
            dcma_b200_run_stats local;
            h.p->run(params->max_iterations, false, mse_history, &local);
            const auto t2 = clk::now();
            h.p->get(DCMA_B200_BUF_DEFORMATION, deformation_out);
            const auto t3 = clk::now();
            if (std::getenv("DCMA_B200_TRACE")) {  // host-side phase times of the one-shot call (diagnostics)
                auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
                std::fprintf(stderr, "[dcma_b200] create %.2f ms, load(H2D) %.2f ms, run %.2f ms (device loop %.2f ms), get(D2H) %.2f ms\n",
                             ms(tc, t0), ms(t0, t1), ms(t1, t2), local.loop_ms, ms(t2, t3));
            }
            if (stats) {
                *stats = local;
                const long long n = geom->slices * geom->rows * geom->cols;
                stats->h2d_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
                stats->d2h_ms = std::chrono::duration<double, std::milli>(t3 - t2).count();
                stats->h2d_bytes = 2 * n * geom->channels * static_cast<long long>(sizeof(float));
                stats->d2h_bytes = 3 * n * static_cast<long long>(sizeof(double));
            }
        },
        errbuf, errbuf_len);
}

int dcma_b200_demons_register_stacks(const dcma_b200_grid *fixed_grid, const float *fixed, const dcma_b200_grid *moving_grid,
                                     const float *moving, const dcma_b200_demons_params *params, double *deformation_out,
                                     double *mse_history, dcma_b200_run_stats *stats, dcma_b200_chunk_fn on_field_chunk,
                                     void *user, char *errbuf, size_t errbuf_len) {
    if (!fixed_grid || !fixed || !moving_grid || !moving || !params || !deformation_out)
        return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    if (stats) std::memset(stats, 0, sizeof(*stats));
    const dcma_b200_geom &geom = fixed_grid->geom;
    const long long nvox = geom.slices * geom.rows * geom.cols;
    const float nan = std::numeric_limits<float>::quiet_NaN();
    if (moving_grid->geom.channels != geom.channels)
        return fail(DCMA_B200_EINVAL, "moving and fixed stacks differ in channel count", errbuf, errbuf_len);
    if (params->pyramid_levels > 1 || params->n_devices > 1) {
        // pyramids and sharded runs take whole host volumes: resample through the host-buffer call, then the one-shot call
        std::vector<float> resampled;
        int rc = guarded([&]() {
            resampled.resize(static_cast<size_t>(nvox * geom.channels));
            dcma::resample_to_grid_host(*moving_grid, moving, *fixed_grid, nan, resampled.data(), params->device);
        }, errbuf, errbuf_len);
        if (rc != DCMA_B200_OK) return rc;
        rc = dcma_b200_demons_register(&geom, fixed, resampled.data(), params, deformation_out, mse_history, stats, errbuf, errbuf_len);
        if (rc == DCMA_B200_OK && on_field_chunk) on_field_chunk(user, 0, nvox);
        return rc;
    }
    return guarded(
        [&]() {
            using clk = std::chrono::steady_clock;
            struct Holder {  // temporaries come from (and go back to) the device pool: no driver call per registration
                dcma::EngineBase *p = nullptr;
                int dev = 0;
                float *d[3] = {nullptr, nullptr, nullptr};
                size_t bytes[3] = {0, 0, 0};
                void drop() {
                    for (int i = 0; i < 3; ++i) {
                        dcma::DevicePool::instance().release(dev, d[i], bytes[i]);
                        d[i] = nullptr;
                    }
                }
                ~Holder() {
                    drop();
                    delete p;
                }
            } h;
            const auto tc = clk::now();
            h.p = make(geom, *params, nullptr);  // selects the device
            cudaStream_t st = static_cast<cudaStream_t>(h.p->stream_handle());
            const long long nmov = moving_grid->geom.slices * moving_grid->geom.rows * moving_grid->geom.cols * geom.channels;
            const size_t bf = sizeof(float) * static_cast<size_t>(nvox * geom.channels), bm = sizeof(float) * static_cast<size_t>(nmov);
            const auto t0 = clk::now();
            h.dev = h.p->device();
            const size_t want[3] = {bf, bm, bf};
            for (int i = 0; i < 3; ++i) {
                void *q = nullptr;
                DCMA_CUDA(dcma::DevicePool::instance().acquire(h.dev, want[i], &q));
                h.d[i] = static_cast<float *>(q);
                h.bytes[i] = want[i];
            }
            DCMA_CUDA(cudaMemcpyAsync(h.d[0], fixed, bf, cudaMemcpyHostToDevice, st));
            DCMA_CUDA(cudaMemcpyAsync(h.d[1], moving, bm, cudaMemcpyHostToDevice, st));
            dcma::resample_to_grid_device(*moving_grid, h.d[1], *fixed_grid, nan, h.d[2], st);   // = cc:964 on the device
            h.p->load(h.d[0], h.d[2], true);
            h.p->sync();
            h.drop();
            const auto t1 = clk::now();
            dcma_b200_run_stats local;
            h.p->run(params->max_iterations, false, mse_history, &local);
            const auto t2 = clk::now();
            h.p->get_deformation_chunked(deformation_out, on_field_chunk, user);
            const auto t3 = clk::now();
            if (std::getenv("DCMA_B200_TRACE")) {
                auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
                std::fprintf(stderr, "[dcma_b200] create %.2f ms, H2D + resample + load %.2f ms, run %.2f ms (device loop %.2f ms), get(D2H, chunked) %.2f ms\n",
                             ms(tc, t0), ms(t0, t1), ms(t1, t2), local.loop_ms, ms(t2, t3));
            }
            if (stats) {
                *stats = local;
                stats->h2d_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
                stats->d2h_ms = std::chrono::duration<double, std::milli>(t3 - t2).count();
                stats->h2d_bytes = static_cast<long long>(bf + bm);
                stats->d2h_bytes = 3 * nvox * static_cast<long long>(sizeof(double));
            }
        },
        errbuf, errbuf_len);
}

int dcma_b200_slab_partition(int64_t slices, int world, int rank, int64_t halo_planes, int64_t *zoff, int64_t *local,
                             int64_t *own0, int64_t *own1) {
    return guarded([&]() {
        const dcma::SlabSpec s = dcma::slab_for_rank(slices, world, rank, halo_planes);
        if (zoff) *zoff = s.zoff;
        if (local) *local = s.local;
        if (own0) *own0 = s.own0;
        if (own1) *own1 = s.own1;
    });
}

int dcma_b200_slab_group_create_local(dcma_b200_slab_group **out, const dcma_b200_geom *geom, const dcma_b200_demons_params *params,
                                      const int32_t *devices, int32_t n_devices, int64_t halo_planes, char *errbuf,
                                      size_t errbuf_len) {
    if (!out || !geom || !params || !devices || n_devices < 1) return fail(DCMA_B200_EINVAL, "bad argument", errbuf, errbuf_len);
    *out = nullptr;
    return guarded(
        [&]() {
            std::vector<int> dev(devices, devices + n_devices);
            *out = new dcma_b200_slab_group{new dcma::SlabGroup(*geom, *params, dev, halo_planes)};
        },
        errbuf, errbuf_len);
}

int dcma_b200_nccl_unique_id(unsigned char id[128], char *errbuf, size_t errbuf_len) {
    if (!id) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    return guarded([&]() { dcma::nccl_unique_id(id); }, errbuf, errbuf_len);
}

int dcma_b200_slab_group_create_nccl(dcma_b200_slab_group **out, const dcma_b200_geom *geom, const dcma_b200_demons_params *params,
                                     const unsigned char id[128], int32_t rank, int32_t world, int32_t device,
                                     int64_t halo_planes, char *errbuf, size_t errbuf_len) {
    if (!out || !geom || !params || !id) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    *out = nullptr;
    return guarded(
        [&]() { *out = new dcma_b200_slab_group{new dcma::SlabGroup(*geom, *params, id, rank, world, device, halo_planes)}; },
        errbuf, errbuf_len);
}

void dcma_b200_slab_group_destroy(dcma_b200_slab_group *g) {
    if (!g) return;
    try {
        delete g->impl;
    } catch (...) {
    }
    delete g;
}

int dcma_b200_slab_group_load(dcma_b200_slab_group *g, const float *fixed, const float *moving, int on_device) {
    if (!g || !fixed || !moving) return fail(DCMA_B200_EINVAL, "null argument", nullptr, 0);
    return guarded([&]() { g->impl->load(fixed, moving, on_device != 0); });
}

int dcma_b200_slab_group_run(dcma_b200_slab_group *g, int64_t max_iterations, double *mse_history, dcma_b200_run_stats *stats) {
    if (!g) return fail(DCMA_B200_EINVAL, "null group", nullptr, 0);
    return guarded([&]() { g->impl->run(max_iterations, mse_history, stats); });
}

int dcma_b200_slab_group_range(dcma_b200_slab_group *g, int64_t *z0, int64_t *z1) {
    if (!g || !z0 || !z1) return fail(DCMA_B200_EINVAL, "null argument", nullptr, 0);
    return guarded([&]() { g->impl->owned_range(z0, z1); });
}

int dcma_b200_slab_group_get_deformation(dcma_b200_slab_group *g, double *deformation_out) {
    if (!g || !deformation_out) return fail(DCMA_B200_EINVAL, "null argument", nullptr, 0);
    return guarded([&]() { g->impl->get_deformation(deformation_out); });
}

int dcma_b200_host_alloc(void **out, size_t bytes, char *errbuf, size_t errbuf_len) {
    if (!out) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    *out = nullptr;
    return guarded(
        [&]() {
            int ndev = 0;
            if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
                throw dcma::CudaError(cudaErrorNoDevice, "no CUDA device is available (dcma_b200 has no CPU fallback)");
            void *p = nullptr;
            DCMA_CUDA(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable));
            *out = p;
        },
        errbuf, errbuf_len);
}

void dcma_b200_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int64_t dcma_b200_slab_group_halo_bytes(const dcma_b200_slab_group *g) { return g ? g->impl->halo_bytes_last_run() : 0; }

int dcma_b200_slab_group_exchange_stats(const dcma_b200_slab_group *g, double *exchange_ms, int64_t *exchanges, int32_t *overlapped,
                                        int32_t *transport, int64_t *halo_planes) {
    if (!g) return fail(DCMA_B200_EINVAL, "null group", nullptr, 0);
    if (exchange_ms) *exchange_ms = g->impl->exchange_ms_last_run();
    if (exchanges) *exchanges = g->impl->exchanges_last_run();
    if (overlapped) *overlapped = g->impl->overlapped_last_run() ? 1 : 0;
    if (transport) *transport = g->impl->transport();
    if (halo_planes) *halo_planes = g->impl->halo();
    return DCMA_B200_OK;
}

int dcma_b200_restrict_image(const dcma_b200_geom *fine, const float *image, dcma_b200_geom *coarse, float *out, int device,
                             char *errbuf, size_t errbuf_len) {
    if (!fine || !coarse || (out && !image)) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    return guarded([&]() { dcma::restrict_image_host(*fine, image, coarse, out, device); }, errbuf, errbuf_len);
}

int dcma_b200_prolong_field(const dcma_b200_geom *fine, const double *coarse_field, double *out, int device, char *errbuf,
                            size_t errbuf_len) {
    if (!fine || !coarse_field || !out) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    return guarded([&]() { dcma::prolong_field_host(*fine, coarse_field, out, device); }, errbuf, errbuf_len);
}

int dcma_b200_warp_image(const dcma_b200_geom *geom, const float *image, const double *deformation, float oob_value,
                         int use_oob_value, float *out, int device, char *errbuf, size_t errbuf_len) {
    if (!geom || !image || !deformation || !out) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    return guarded(
        [&]() { dcma::warp_image_host(*geom, image, deformation, use_oob_value != 0, oob_value, out, device); }, errbuf,
        errbuf_len);
}

int dcma_b200_histogram_match(const float *source, int64_t n_source, const float *reference, int64_t n_reference,
                              int64_t num_bins, double outlier_fraction, float *out, int32_t *mapped, int device, char *errbuf,
                              size_t errbuf_len) {
    if (!source || !reference || !out) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    return guarded(
        [&]() {
            const int m = dcma::histogram_match_host(source, n_source, reference, n_reference, num_bins, outlier_fraction, out, device);
            if (mapped) *mapped = m;
        },
        errbuf, errbuf_len);
}

int dcma_b200_warp_image_grid(const dcma_b200_grid *field_grid, const double *deformation, const dcma_b200_grid *image_grid,
                              const float *image, const unsigned char *mask, float oob_value, float *out, int device,
                              char *errbuf, size_t errbuf_len) {
    if (!field_grid || !deformation || !image_grid || !image || !out) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    return guarded(
        [&]() { dcma::warp_image_grid_host(*field_grid, deformation, *image_grid, image, mask, oob_value, out, device); }, errbuf,
        errbuf_len);
}

int dcma_b200_warp_to_grid(const dcma_b200_grid *field_grid, const double *deformation, const dcma_b200_grid *source_grid,
                           const float *source, const dcma_b200_grid *out_grid, const unsigned char *mask, int32_t channel,
                           float oob_value, float *out_inout, int device, char *errbuf, size_t errbuf_len) {
    if (!field_grid || !deformation || !source_grid || !source || !out_grid || !out_inout)
        return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    return guarded(
        [&]() { dcma::warp_to_grid_host(*field_grid, deformation, *source_grid, source, *out_grid, mask, channel, oob_value, out_inout, device); },
        errbuf, errbuf_len);
}

int dcma_b200_virtual_field(int32_t kind, const dcma_b200_grid *grid, const double *bounds, double *field_out, double *extremes,
                            int device, char *errbuf, size_t errbuf_len) {
    if (!grid || !bounds || !field_out) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    return guarded([&]() { dcma::virtual_field_host(kind, *grid, bounds, field_out, extremes, device); }, errbuf, errbuf_len);
}

int dcma_b200_resample_to_grid(const dcma_b200_grid *src_grid, const float *image, const dcma_b200_grid *dst_grid, float oob_value,
                               float *out, int device, char *errbuf, size_t errbuf_len) {
    if (!src_grid || !image || !dst_grid || !out) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    return guarded([&]() { dcma::resample_to_grid_host(*src_grid, image, *dst_grid, oob_value, out, device); }, errbuf, errbuf_len);
}

int dcma_b200_format_doubles(const double *values, int64_t n, char *text, int64_t capacity, int64_t *text_len, int device,
                             char *errbuf, size_t errbuf_len) {
    if ((n > 0 && (!values || !text)) || !text_len) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    *text_len = 0;
    return guarded([&]() { *text_len = dcma::format_doubles_host(values, n, text, capacity, device); }, errbuf, errbuf_len);
}

int dcma_b200_parse_doubles(const char *text, int64_t text_len, int64_t n, double *values, int64_t *consumed, int device,
                            char *errbuf, size_t errbuf_len) {
    if (n > 0 && (!values || (!text && text_len > 0))) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    if (consumed) *consumed = 0;
    return guarded(
        [&]() {
            const long long used = dcma::parse_doubles_host(text, text_len, n, values, device);
            if (consumed) *consumed = used;
        },
        errbuf, errbuf_len);
}

}  // extern "C"
