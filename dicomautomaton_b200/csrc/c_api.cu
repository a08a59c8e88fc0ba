// dcma_b200 -- the C ABI (include/dcma_b200/demons_c_api.h).  Nothing throws across this boundary.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <exception>
#include <new>
#include <stdexcept>
#include <string>

#include "common.cuh"
#include "engine.h"

struct dcma_b200_engine {
    dcma::EngineBase *impl;
    dcma_b200_geom geom;
};

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
    if (p.pyramid_levels > 1 || p.use_symmetric_force)
        throw std::invalid_argument("pyramid_levels > 1 / use_symmetric_force are not available at the engine level");
    switch (p.field_mode) {
        case DCMA_B200_FIELD_FP64: return dcma::make_engine_f64(g, p, stream);
        case DCMA_B200_FIELD_FP64_STRICT: return dcma::make_engine_f64_strict(g, p, stream);
        case DCMA_B200_FIELD_FP32: return dcma::make_engine_f32(g, p, stream);
        default: throw std::invalid_argument("unknown field_mode");
    }
}

}  // namespace

extern "C" {

int dcma_b200_abi_version(void) { return DCMA_B200_ABI_VERSION; }

const char *dcma_b200_last_error(void) { return g_last_error.c_str(); }

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
            h.p = make(*geom, *params, nullptr);
            const auto t0 = clk::now();
            h.p->load(fixed, moving, false);
            const auto t1 = clk::now();
            dcma_b200_run_stats local;
            h.p->run(params->max_iterations, false, mse_history, &local);
            const auto t2 = clk::now();
            h.p->get(DCMA_B200_BUF_DEFORMATION, deformation_out);
            const auto t3 = clk::now();
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

int dcma_b200_warp_image(const dcma_b200_geom *geom, const float *image, const double *deformation, float oob_value,
                         int use_oob_value, float *out, int device, char *errbuf, size_t errbuf_len) {
    if (!geom || !image || !deformation || !out) return fail(DCMA_B200_EINVAL, "null argument", errbuf, errbuf_len);
    return guarded(
        [&]() { dcma::warp_image_host(*geom, image, deformation, use_oob_value != 0, oob_value, out, device); }, errbuf,
        errbuf_len);
}

}  // extern "C"
