// dcma_b200 -- precision-mode independent engine interface (host side).  The C ABI (c_api.cu) talks to this only.
#pragma once
#include <cstdint>

#include "../../include/dcma_b200/demons_c_api.h"

namespace dcma {

// Device-resident replacement of the reference's `sycl_demons_engine` (src/Alignment_Demons.cc:52-704).
class EngineBase {
  public:
    virtual ~EngineBase() {}
    virtual void load(const float *fixed, const float *moving, bool on_device) = 0;
    virtual void run(int64_t max_iterations, bool profile, double *mse_history_host, dcma_b200_run_stats *stats) = 0;
    virtual void stage(int stage, double sigma_mm, double *mse_out, int64_t *n_valid_out) = 0;
    virtual void get(int buffer, void *host) = 0;
    virtual void set(int buffer, const void *host) = 0;
    virtual void export_deformation_device(double *device_out) = 0;
    virtual void sync() = 0;
    virtual int64_t device_bytes() const = 0;
    virtual int device() const = 0;
};

// one factory per precision-mode translation unit
EngineBase *make_engine_f64(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream);
EngineBase *make_engine_f64_strict(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream);
EngineBase *make_engine_f32(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream);

// warp_moving semantics on host buffers (strict fp64 translation unit)
void warp_image_host(const dcma_b200_geom &g, const float *image, const double *deformation_aos, bool use_oob,
                     float oob_value, float *out, int device);

}  // namespace dcma
