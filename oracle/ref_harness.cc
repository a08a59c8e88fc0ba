// TEST INFRASTRUCTURE ONLY (oracle/).  C-ABI harness around the UNMODIFIED reference Demons engine.
//
// The engine class `sycl_demons_engine` (/root/reference/src/Alignment_Demons.cc:52-704) is file-local in a
// translation unit that needs Ygor, so it cannot be compiled as a file.  `oracle/build_ref.py` extracts, by textual
// anchors and only into the git-ignored `oracle/_ref/`, (i) the parameter/volume declarations
// (/root/reference/src/Alignment_Demons.h:19-79) and (ii) the engine class, and this harness #includes those two
// generated fragments together with the reference's own /root/reference/src/SYCL_Fallback.h (included in place).
// Nothing of the reference is stored in the repository.
//
// What is restated here (our code): the driver loop of AlignViaDemons (/root/reference/src/Alignment_Demons.cc:980-995)
// because the surrounding function needs Ygor images; it is 10 lines and is followed literally.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the resulting
// oracle/_ref/libdemons_ref.so.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <exception>
#include <functional>
#include <limits>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#define YLOGINFO(x) do { } while (0)
#define YLOGWARN(x) do { } while (0)

#include "_ref/ref_params_extract.inc"  // AlignViaDemonsParams, demons_volume<T>, vol_idx  (generated)

#include "SYCL_Fallback.h"              // the reference's own file, found via -I/root/reference/src
namespace dcma_sycl = sycl;

#define private public                  // expose per-stage methods and buffers for stage-level parity
#include "_ref/ref_engine_extract.inc"  // class sycl_demons_engine  (generated)
#undef private

extern "C" {

typedef struct {
    int64_t slices, rows, cols, channels;
    double pxl_dx, pxl_dy, pxl_dz;
} ref_geom;

typedef struct {
    int64_t max_iterations;
    double convergence_threshold;
    double deformation_field_smoothing_sigma;
    double update_field_smoothing_sigma;
    int32_t use_diffeomorphic;
    int32_t _pad;
    double normalization_factor;
    double max_update_magnitude;
} ref_params;

}  // extern "C"

namespace {

AlignViaDemonsParams to_params(const ref_params *p) {
    AlignViaDemonsParams q;
    q.max_iterations = p->max_iterations;
    q.convergence_threshold = p->convergence_threshold;
    q.deformation_field_smoothing_sigma = p->deformation_field_smoothing_sigma;
    q.update_field_smoothing_sigma = p->update_field_smoothing_sigma;
    q.use_diffeomorphic = (p->use_diffeomorphic != 0);
    q.normalization_factor = p->normalization_factor;
    q.max_update_magnitude = p->max_update_magnitude;
    q.verbosity = 0;
    return q;
}

demons_volume<float> to_volume(const ref_geom *g, const float *data) {
    demons_volume<float> v;
    v.slices = g->slices;
    v.rows = g->rows;
    v.cols = g->cols;
    v.channels = g->channels;
    v.pxl_dx = g->pxl_dx;
    v.pxl_dy = g->pxl_dy;
    v.pxl_dz = g->pxl_dz;
    const size_t n = static_cast<size_t>(g->slices * g->rows * g->cols * g->channels);
    v.data.assign(data, data + n);
    return v;
}

}  // namespace

extern "C" {

unsigned ref_worker_threads(void) {
    const unsigned hw = std::thread::hardware_concurrency();
    return (hw == 0U) ? 2U : hw;  // SYCL_Fallback.h:350-355
}

void *ref_engine_create(const ref_geom *g, const float *fixed, const float *moving, const ref_params *p) {
    try {
        return new sycl_demons_engine(to_volume(g, fixed), to_volume(g, moving), to_params(p));
    } catch (...) {
        return nullptr;
    }
}

void ref_engine_destroy(void *e) { delete static_cast<sycl_demons_engine *>(e); }

double ref_engine_iterate(void *e) { return static_cast<sycl_demons_engine *>(e)->compute_single_iteration(); }

void ref_engine_compute_update_and_mse(void *e, double *mse, int64_t *n) {
    static_cast<sycl_demons_engine *>(e)->compute_update_and_mse(*mse, *n);
}
void ref_engine_smooth_update(void *e, double sigma_mm) {
    auto *E = static_cast<sycl_demons_engine *>(e);
    E->smooth_on_device(E->update_dev, sigma_mm);
}
void ref_engine_smooth_deformation(void *e, double sigma_mm) {
    auto *E = static_cast<sycl_demons_engine *>(e);
    E->smooth_on_device(E->deformation_dev, sigma_mm);
}
void ref_engine_add_update(void *e) { static_cast<sycl_demons_engine *>(e)->add_update(); }
void ref_engine_add_update_diffeomorphic(void *e) { static_cast<sycl_demons_engine *>(e)->add_update_diffeomorphic(); }
void ref_engine_warp_moving(void *e) { static_cast<sycl_demons_engine *>(e)->warp_moving(); }

// which: 0 fixed, 1 moving, 2 warped (float[N*C]); 3 gradient, 4 update, 5 deformation (double[N*3])
static void *buf_of(sycl_demons_engine *E, int which, size_t *bytes) {
    const size_t sb = E->scalar_volume_size * sizeof(float);
    const size_t vb = E->vector_volume_size * sizeof(double);
    switch (which) {
        case 0: *bytes = sb; return E->fixed_dev;
        case 1: *bytes = sb; return E->moving_dev;
        case 2: *bytes = sb; return E->warped_dev;
        case 3: *bytes = vb; return E->gradient_dev;
        case 4: *bytes = vb; return E->update_dev;
        case 5: *bytes = vb; return E->deformation_dev;
        default: *bytes = 0; return nullptr;
    }
}
int ref_engine_get(void *e, int which, void *out) {
    size_t bytes = 0;
    void *p = buf_of(static_cast<sycl_demons_engine *>(e), which, &bytes);
    if (!p) return 1;
    std::memcpy(out, p, bytes);
    return 0;
}
int ref_engine_set(void *e, int which, const void *in) {
    size_t bytes = 0;
    void *p = buf_of(static_cast<sycl_demons_engine *>(e), which, &bytes);
    if (!p) return 1;
    std::memcpy(p, in, bytes);
    return 0;
}

// Driver loop, restating /root/reference/src/Alignment_Demons.cc:977-997 on raw volumes.
// mse_history must hold max_iterations doubles.  Returns 0 on success.
int ref_demons_run(const ref_geom *g, const float *fixed, const float *moving, const ref_params *p,
                   double *deformation_out, double *mse_history, int64_t *iterations_done) {
    try {
        sycl_demons_engine engine(to_volume(g, fixed), to_volume(g, moving), to_params(p));
        double prev_mse = std::numeric_limits<double>::infinity();
        int64_t done = 0;
        for (int64_t iter = 0; iter < p->max_iterations; ++iter) {
            const double mse = engine.compute_single_iteration();
            if (mse_history) mse_history[iter] = mse;
            done = iter + 1;
            const double mse_change = std::abs(prev_mse - mse);
            if (mse_change < p->convergence_threshold && iter > 0) {
                break;
            }
            prev_mse = mse;
        }
        if (iterations_done) *iterations_done = done;
        auto vol = engine.export_deformation_volume();
        std::copy(vol.data.begin(), vol.data.end(), deformation_out);
        return 0;
    } catch (...) {
        return 1;
    }
}

}  // extern "C"
