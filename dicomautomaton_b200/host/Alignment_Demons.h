// Alignment_Demons.h -- drop-in replacement for DICOMautomaton's src/Alignment_Demons.h whose iteration loop runs on a
// B200 through the dcma_b200 C ABI (include/dcma_b200/demons_c_api.h) instead of the SYCL engine.
//
// Same public surface as the reference header (names, signatures, defaults, error behaviour):
//   AlignViaDemonsParams                      reference :20-61   (+ additive extension fields, reference-preserving defaults)
//   demons_volume<T>, vol_idx                 reference :64-79
//   AlignViaDemonsHelpers::resample_image_to_reference_grid / histogram_match / get_image /
//       warp_image_with_field / marshal_collection_to_volume / marshal_volume_to_collection    reference :82-206
//   AlignViaDemons                            reference :233-236
// Install: replace src/Alignment_Demons.{h,cc} with these two files and link `dcma_b200`.
#pragma once

#include <cstdint>
#include <functional>
#include <iosfwd>
#include <list>
#include <optional>
#include <stdexcept>
#include <vector>

#include "YgorImages.h"
#include "YgorLog.h"
#include "YgorMath.h"
#include "YgorMisc.h"

#include "Alignment_Field.h"

// Controls of the Demons registration; field-for-field the reference struct, then extensions.
struct AlignViaDemonsParams {
    int64_t max_iterations = 100;                    // upper bound on iterations
    double convergence_threshold = 0.001;            // stop when |MSE change| drops below this
    double deformation_field_smoothing_sigma = 1.0;  // mm; Gaussian regularisation of the deformation field
    double update_field_smoothing_sigma = 0.5;       // mm; Gaussian smoothing of the update (diffeomorphic only)
    bool use_diffeomorphic = false;                  // compose updates instead of adding them
    bool use_histogram_matching = false;             // match the moving image's intensity distribution first
    int64_t histogram_bins = 256;
    double histogram_outlier_fraction = 0.01;        // e.g. 0.01 -> 1st..99th percentile
    double normalization_factor = 1.0;               // alpha in the Demons force denominator
    double max_update_magnitude = 2.0;               // clamp on the per-iteration update
    int64_t verbosity = 1;

    // ---- dcma_b200 extensions (absent in the reference; defaults reproduce the reference) ----
    int32_t field_mode = 0;           // 0: fp64 fields, 1: fp64 bit-exact vs the reference's CPU build, 2: fp32 fields, 3: fp64 D + fp32 U
    int32_t device = -1;              // CUDA device ordinal, -1 = current
    int32_t pyramid_levels = 1;       // multi-resolution levels (1 = single resolution, as the reference)
    bool use_symmetric_force = false; // (grad F + grad W)/2 in the force (false = fixed-image gradient, as the reference)
    int32_t pyramid_iterations[3] = {0, 0, 0};  // iterations at pyramid levels 1..3 (0 = max_iterations)
    int32_t n_devices = 0;            // > 1: shard the volume into z-slabs over GPUs 0..n_devices-1 of this process
    int32_t halo_planes = 0;          // z-slab halo (0 = default)
};

// Dense proxy of a regular image stack (reference :64-74).
template <class T>
struct demons_volume {
    int64_t slices = 0;
    int64_t rows = 0;
    int64_t cols = 0;
    int64_t channels = 0;
    double pxl_dx = 1.0;
    double pxl_dy = 1.0;
    double pxl_dz = 1.0;
    std::vector<T> data;
};

template <class T>
inline size_t vol_idx(const demons_volume<T> &v, int64_t z, int64_t y, int64_t x, int64_t c) {
    return static_cast<size_t>((((z * v.rows + y) * v.cols + x) * v.channels) + c);
}

namespace AlignViaDemonsHelpers {

// Sample `moving` at every voxel position of `reference` (trilinear); voxels that cannot be sampled become NaN.
planar_image_collection<float, double> resample_image_to_reference_grid(
    const planar_image_collection<float, double> &moving, const planar_image_collection<float, double> &reference);

// Map the intensity distribution of `source` onto that of `reference` (binned CDF matching between percentile bounds).
planar_image_collection<float, double> histogram_match(const planar_image_collection<float, double> &source,
                                                       const planar_image_collection<float, double> &reference,
                                                       int64_t num_bins, double outlier_fraction);

// Random access into a std::list of images (O(i); prefer iterating).
template <class T, class R>
planar_image<T, R> &get_image(std::list<planar_image<T, R>> &imgs, size_t i) {
    if (!isininc(0, i, imgs.size())) throw std::runtime_error("Requested image index not present, unable to continue");
    return *std::next(std::begin(imgs), i);
}
template <class T, class R>
const planar_image<T, R> &get_image(const std::list<planar_image<T, R>> &imgs, size_t i) {
    if (!isininc(0, i, imgs.size())) throw std::runtime_error("Requested image index not present, unable to continue");
    return *std::next(std::begin(imgs), i);
}

// Pull-warp an image stack through a deformation field: out(p) = img(field.transform(p)).
planar_image_collection<float, double> warp_image_with_field(const planar_image_collection<float, double> &img_coll,
                                                             const deformation_field &def_field);

// planar_image_collection (Drover side) -> dense volume (device side).  Slice order = list order.
template <class T>
demons_volume<T> marshal_collection_to_volume(const planar_image_collection<T, double> &coll) {
    if (coll.images.empty()) throw std::invalid_argument("Cannot marshal empty image collection");
    auto &nc = const_cast<planar_image_collection<T, double> &>(coll);
    std::list<std::reference_wrapper<planar_image<T, double>>> refs;
    for (auto &img : nc.images) refs.push_back(std::ref(img));
    if (!Images_Form_Regular_Grid(refs))
        throw std::invalid_argument("Image collection is not a regular rectilinear grid and cannot be marshaled as a dense volume");
    const auto &first = coll.images.front();
    demons_volume<T> v;
    v.slices = static_cast<int64_t>(coll.images.size());
    v.rows = first.rows;
    v.cols = first.columns;
    v.channels = first.channels;
    v.pxl_dx = first.pxl_dx;
    v.pxl_dy = first.pxl_dy;
    v.pxl_dz = first.pxl_dz;
    v.data.resize(static_cast<size_t>(v.slices * v.rows * v.cols * v.channels), T{});
    int64_t z = 0;
    for (const auto &img : coll.images) {
        if (img.rows != v.rows || img.columns != v.cols || img.channels != v.channels)
            throw std::invalid_argument("Image collection has inconsistent rows/columns/channels and cannot be marshaled as a rectilinear volume");
        for (int64_t y = 0; y < v.rows; ++y)
            for (int64_t x = 0; x < v.cols; ++x)
                for (int64_t c = 0; c < v.channels; ++c) v.data[vol_idx(v, z, y, x, c)] = static_cast<T>(img.value(y, x, c));
        ++z;
    }
    return v;
}

// Dense volume -> one image per slice carrying the geometry and metadata of the matching reference slice.
// (One pass over the list: the reference walks it from the front for every slice.)
template <class T>
planar_image_collection<T, double> marshal_volume_to_collection(const demons_volume<T> &vol,
                                                                const planar_image_collection<float, double> &reference_geometry) {
    if (static_cast<int64_t>(reference_geometry.images.size()) < vol.slices)
        throw std::runtime_error("Requested image index not present, unable to continue");
    planar_image_collection<T, double> out;
    auto it = reference_geometry.images.begin();
    for (int64_t z = 0; z < vol.slices; ++z, ++it) {
        const auto &ref_img = *it;
        out.images.emplace_back();
        auto &img = out.images.back();
        img.init_orientation(ref_img.row_unit, ref_img.col_unit);
        img.init_buffer(vol.rows, vol.cols, vol.channels);
        img.init_spatial(ref_img.pxl_dx, ref_img.pxl_dy, ref_img.pxl_dz, ref_img.anchor, ref_img.offset);
        img.metadata = ref_img.metadata;
        for (int64_t y = 0; y < vol.rows; ++y)
            for (int64_t x = 0; x < vol.cols; ++x)
                for (int64_t c = 0; c < vol.channels; ++c) img.reference(y, x, c) = vol.data[vol_idx(vol, z, y, x, c)];
    }
    return out;
}

}  // namespace AlignViaDemonsHelpers

// Deformable registration of `moving` onto `stationary` with the Demons algorithm.  Returns the pull deformation field
// (warped(p) = moving(p + D(p)), mm) on the stationary grid, or std::nullopt after logging a warning on any failure.
std::optional<deformation_field> AlignViaDemons(AlignViaDemonsParams &params,
                                                const planar_image_collection<float, double> &moving,
                                                const planar_image_collection<float, double> &stationary);
