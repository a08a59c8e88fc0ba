// dropin_harness.cc -- a C entry around the C++ drop-in `AlignViaDemons` so that bench.py (ctypes) can time the call a
// DICOMautomaton operation makes (src/Operations/RegisterImagesDemons.cc:337): image stacks in, deformation_field out,
// BOTH marshalling steps, the grid resampling and the pinned-buffer handling inside the timed region.  The image stacks
// themselves are the caller's data (the Drover's Image_Arrays) and are built once, outside it.
// Built against the minimal Ygor stand-in (host/shim) in this repository; inside the reference tree the same
// Alignment_Demons.cc compiles against the real Ygor.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>

#include "Alignment_Demons.h"

namespace {
struct Harness {
    planar_image_collection<float, double> fixed, moving;
};
planar_image_collection<float, double> make_stack(int64_t slices, int64_t rows, int64_t cols, const float *data, double dx, double dy,
                                                  double dz) {
    planar_image_collection<float, double> c;
    const size_t per = static_cast<size_t>(rows * cols);
    for (int64_t z = 0; z < slices; ++z) {
        c.images.emplace_back();
        auto &img = c.images.back();
        img.init_orientation(vec3<double>(1, 0, 0), vec3<double>(0, 1, 0));
        img.init_buffer(rows, cols, 1);
        img.init_spatial(dx, dy, dz, vec3<double>(0, 0, 0), vec3<double>(0, 0, dz * static_cast<double>(z)));
        for (int64_t y = 0; y < rows; ++y)
            for (int64_t x = 0; x < cols; ++x) img.reference(y, x, 0) = data[static_cast<size_t>(z) * per + static_cast<size_t>(y * cols + x)];
        img.metadata["SliceNumber"] = std::to_string(z);
    }
    return c;
}
}  // namespace

extern "C" {

__attribute__((visibility("default"))) void *dcma_dropin_create(int64_t slices, int64_t rows, int64_t cols, const float *fixed,
                                                                const float *moving, double dx, double dy, double dz) {
    try {
        auto h = std::make_unique<Harness>();
        h->fixed = make_stack(slices, rows, cols, fixed, dx, dy, dz);
        h->moving = make_stack(slices, rows, cols, moving, dx, dy, dz);
        return h.release();
    } catch (...) {
        return nullptr;
    }
}

__attribute__((visibility("default"))) void dcma_dropin_destroy(void *h) { delete static_cast<Harness *>(h); }

// One AlignViaDemons call.  *seconds: wall time of the call alone.  out3: the displacement of the voxel `probe_voxel`
// (so that the caller can check the result against the C-ABI path).  Returns 0 on success, 1 if it returned nullopt.
__attribute__((visibility("default"))) int dcma_dropin_run(void *hv, int64_t max_iterations, double threshold, double sigma_d,
                                                           double sigma_u, int diffeomorphic, int field_mode, int device,
                                                           int64_t probe_voxel, double *seconds, double *out3) {
    auto *h = static_cast<Harness *>(hv);
    AlignViaDemonsParams p;
    p.max_iterations = max_iterations;
    p.convergence_threshold = threshold;
    p.deformation_field_smoothing_sigma = sigma_d;
    p.update_field_smoothing_sigma = sigma_u;
    p.use_diffeomorphic = diffeomorphic != 0;
    p.verbosity = 0;
    p.field_mode = field_mode;
    p.device = device;
    const auto t0 = std::chrono::steady_clock::now();
    auto res = AlignViaDemons(p, h->moving, h->fixed);
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (!res) return 1;
    if (out3) {
        const auto &coll = res->get_imagecoll_crefw().get();
        const auto &first = coll.images.front();
        const int64_t per = first.rows * first.columns;
        const int64_t z = probe_voxel / per, r = (probe_voxel % per) / first.columns, c = probe_voxel % first.columns;
        auto it = coll.images.begin();
        std::advance(it, z);
        for (int k = 0; k < 3; ++k) out3[k] = it->value(r, c, k);
    }
    return 0;
}
}
