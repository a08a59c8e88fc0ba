// Alignment_Demons.cc -- host side of the B200 Demons drop-in (replaces DICOMautomaton's src/Alignment_Demons.cc).
//
// What stays on the host, exactly as in the reference: resampling onto the stationary grid (reference cc:709-755),
// optional histogram matching (cc:760-891), marshalling between image stacks and dense volumes, building the
// deformation_field, and the error contract (empty input / any exception -> warning + nullopt, cc:953-956, 1001-1004).
// What moves to the GPU: the engine construction, the whole iteration loop and the export of the field (cc:977-997),
// through ONE blocking C-ABI call, dcma_b200_demons_register().  There is no CPU fallback: if the library reports an
// error (no device, out of memory, ...) the registration fails the way the reference's SYCL exceptions do.
#include "Alignment_Demons.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>

#include "dcma_b200/demons_c_api.h"

using namespace AlignViaDemonsHelpers;

planar_image_collection<float, double> AlignViaDemonsHelpers::resample_image_to_reference_grid(
    const planar_image_collection<float, double> &moving, const planar_image_collection<float, double> &reference) {
    if (moving.images.empty() || reference.images.empty())
        throw std::invalid_argument("Cannot resample: image collection is empty");
    const float nan = std::numeric_limits<float>::quiet_NaN();
    planar_image_collection<float, double> out;
    for (const auto &ref_img : reference.images) {
        out.images.push_back(ref_img);  // geometry and metadata of the reference slice
        auto &img = out.images.back();
        std::fill(img.data.begin(), img.data.end(), nan);
        for (int64_t r = 0; r < img.rows; ++r)
            for (int64_t c = 0; c < img.columns; ++c) {
                const auto p = ref_img.position(r, c);
                for (int64_t ch = 0; ch < img.channels; ++ch) {
                    const auto v = moving.trilinearly_interpolate(p, ch, nan);
                    if (std::isfinite(v)) img.reference(r, c, ch) = v;
                }
            }
    }
    return out;
}

namespace {
std::vector<double> finite_sorted(const planar_image_collection<float, double> &coll) {
    std::vector<double> v;
    for (const auto &img : coll.images)
        for (const auto &x : img.data)
            if (std::isfinite(x)) v.push_back(x);
    std::sort(v.begin(), v.end());
    return v;
}
double at_fraction(const std::vector<double> &sorted, double f) {
    return sorted[static_cast<size_t>(f * static_cast<double>(sorted.size() - 1))];
}
// Normalised cumulative histogram of the values inside [lo, hi].
std::vector<double> cdf_between(const std::vector<double> &vals, double lo, double hi, int64_t bins) {
    std::vector<double> h(static_cast<size_t>(bins), 0.0);
    for (double v : vals)
        if (v >= lo && v <= hi) {
            const int64_t b = std::min<int64_t>(bins - 1, static_cast<int64_t>((v - lo) / (hi - lo) * static_cast<double>(bins)));
            h[static_cast<size_t>(b)] += 1.0;
        }
    for (int64_t i = 1; i < bins; ++i) h[static_cast<size_t>(i)] += h[static_cast<size_t>(i - 1)];
    const double total = h.back();
    if (total > 0.0)
        for (auto &x : h) x /= total;
    return h;
}
}  // namespace

planar_image_collection<float, double> AlignViaDemonsHelpers::histogram_match(
    const planar_image_collection<float, double> &source, const planar_image_collection<float, double> &reference,
    int64_t num_bins, double outlier_fraction) {
    if (source.images.empty() || reference.images.empty())
        throw std::invalid_argument("Cannot perform histogram matching: image collection is empty");
    const auto src = finite_sorted(source), ref = finite_sorted(reference);
    if (src.empty() || ref.empty()) {
        YLOGWARN("No valid pixel values found for histogram matching, returning source unchanged");
        return source;
    }
    const double src_lo = at_fraction(src, outlier_fraction), src_hi = at_fraction(src, 1.0 - outlier_fraction);
    const double ref_lo = at_fraction(ref, outlier_fraction), ref_hi = at_fraction(ref, 1.0 - outlier_fraction);
    if (src_hi <= src_lo || ref_hi <= ref_lo) {
        YLOGWARN("Image has constant or near-constant intensity, histogram matching not applicable");
        return source;
    }
    const auto src_cdf = cdf_between(src, src_lo, src_hi, num_bins);
    const auto ref_cdf = cdf_between(ref, ref_lo, ref_hi, num_bins);
    // source bin -> first reference bin whose CDF reaches the source bin's quantile -> reference intensity
    std::vector<double> lut(static_cast<size_t>(num_bins));
    for (int64_t b = 0; b < num_bins; ++b) {
        const auto it = std::find_if(ref_cdf.begin(), ref_cdf.end(), [&](double c) { return c >= src_cdf[static_cast<size_t>(b)]; });
        const int64_t rb = (it == ref_cdf.end()) ? 0 : static_cast<int64_t>(it - ref_cdf.begin());
        lut[static_cast<size_t>(b)] = ref_lo + (ref_hi - ref_lo) * static_cast<double>(rb) / static_cast<double>(num_bins);
    }
    auto out = source;
    for (auto &img : out.images)
        for (auto &v : img.data) {
            if (!std::isfinite(v)) continue;
            if (v < src_lo) {
                v = static_cast<float>(ref_lo);
            } else if (v > src_hi) {
                v = static_cast<float>(ref_hi);
            } else {
                const int64_t b = std::min<int64_t>(num_bins - 1, static_cast<int64_t>((v - src_lo) / (src_hi - src_lo) * static_cast<double>(num_bins)));
                v = static_cast<float>(lut[static_cast<size_t>(b)]);
            }
        }
    return out;
}

planar_image_collection<float, double> AlignViaDemonsHelpers::warp_image_with_field(
    const planar_image_collection<float, double> &img_coll, const deformation_field &def_field) {
    if (img_coll.images.empty()) throw std::invalid_argument("Cannot warp: image collection is empty");
    // adjacency index over the SOURCE images so that single-slice stacks are still sampled bilinearly
    auto &nc = const_cast<planar_image_collection<float, double> &>(img_coll);
    const planar_image_adjacency<float, double> src({}, {{std::ref(nc)}}, img_coll.images.front().ortho_unit());
    const float nan = std::numeric_limits<float>::quiet_NaN();
    auto out = img_coll;
    for (auto &img : out.images)
        for (int64_t r = 0; r < img.rows; ++r)
            for (int64_t c = 0; c < img.columns; ++c) {
                const auto q = def_field.transform(img.position(r, c));
                for (int64_t ch = 0; ch < img.channels; ++ch) img.reference(r, c, ch) = src.trilinearly_interpolate(q, ch, nan);
            }
    return out;
}

std::optional<deformation_field> AlignViaDemons(AlignViaDemonsParams &params,
                                                const planar_image_collection<float, double> &moving_in,
                                                const planar_image_collection<float, double> &stationary) {
    if (moving_in.images.empty() || stationary.images.empty()) {
        YLOGWARN("Unable to perform demons alignment: an image array is empty");
        return std::nullopt;
    }
    try {
        if (params.verbosity >= 1) YLOGINFO("Resampling moving image to reference grid");
        auto moving = resample_image_to_reference_grid(moving_in, stationary);
        const auto fixed_vol = marshal_collection_to_volume(stationary);
        auto moving_vol = marshal_collection_to_volume(moving);
        if (params.use_histogram_matching) {
            // reference: histogram_match(moving, stationary, ...) on the collections (cc:967-972).  Here the same
            // algorithm runs on the GPU over the dense volumes (bit-identical result; the public helper above stays a
            // host function, as in the reference)
            if (params.verbosity >= 1) YLOGINFO("Applying histogram matching");
            std::vector<float> matched(moving_vol.data.size());
            char herr[512] = {0};
            int32_t mapped = 0;
            if (dcma_b200_histogram_match(moving_vol.data.data(), static_cast<int64_t>(moving_vol.data.size()),
                                          fixed_vol.data.data(), static_cast<int64_t>(fixed_vol.data.size()),
                                          params.histogram_bins, params.histogram_outlier_fraction, matched.data(), &mapped,
                                          params.device, herr, sizeof herr) != DCMA_B200_OK)
                throw std::runtime_error(herr);
            if (!mapped) YLOGWARN("Histogram matching not applicable (no finite values or constant intensity), moving image unchanged");
            moving_vol.data.swap(matched);
        }

        dcma_b200_geom g;
        g.slices = fixed_vol.slices;
        g.rows = fixed_vol.rows;
        g.cols = fixed_vol.cols;
        g.channels = fixed_vol.channels;
        g.pxl_dx = fixed_vol.pxl_dx;
        g.pxl_dy = fixed_vol.pxl_dy;
        g.pxl_dz = fixed_vol.pxl_dz;

        dcma_b200_demons_params p;
        dcma_b200_demons_params_default(&p);
        p.max_iterations = params.max_iterations;
        p.convergence_threshold = params.convergence_threshold;
        p.deformation_field_smoothing_sigma = params.deformation_field_smoothing_sigma;
        p.update_field_smoothing_sigma = params.update_field_smoothing_sigma;
        p.use_diffeomorphic = params.use_diffeomorphic ? 1 : 0;
        p.use_histogram_matching = params.use_histogram_matching ? 1 : 0;
        p.histogram_bins = params.histogram_bins;
        p.histogram_outlier_fraction = params.histogram_outlier_fraction;
        p.normalization_factor = params.normalization_factor;
        p.max_update_magnitude = params.max_update_magnitude;
        p.verbosity = params.verbosity;
        p.field_mode = params.field_mode;
        if (const char *e = std::getenv("DCMA_B200_FIELD_MODE")) {  // test matrix / site policy override
            if (!std::strcmp(e, "fp64")) p.field_mode = DCMA_B200_FIELD_FP64;
            if (!std::strcmp(e, "fp64_strict")) p.field_mode = DCMA_B200_FIELD_FP64_STRICT;
            if (!std::strcmp(e, "fp32")) p.field_mode = DCMA_B200_FIELD_FP32;
        }
        p.device = params.device;
        p.pyramid_levels = params.pyramid_levels;
        p.use_symmetric_force = params.use_symmetric_force ? 1 : 0;
        for (int i = 0; i < 3; ++i) p.pyramid_iterations[i] = params.pyramid_iterations[i];
        p.n_devices = params.n_devices;
        if (const char *e = std::getenv("DCMA_B200_DEVICES")) p.n_devices = std::atoi(e);  // site policy: GPUs to shard over
        p.halo_planes = params.halo_planes;

        demons_volume<double> def;
        def.slices = fixed_vol.slices;
        def.rows = fixed_vol.rows;
        def.cols = fixed_vol.cols;
        def.channels = 3;
        def.pxl_dx = fixed_vol.pxl_dx;
        def.pxl_dy = fixed_vol.pxl_dy;
        def.pxl_dz = fixed_vol.pxl_dz;
        def.data.resize(static_cast<size_t>(def.slices * def.rows * def.cols * 3));

        char err[512] = {0};
        dcma_b200_run_stats stats;
        const int rc = dcma_b200_demons_register(&g, fixed_vol.data.data(), moving_vol.data.data(), &p, def.data.data(),
                                                 nullptr, &stats, err, sizeof(err));
        if (rc != DCMA_B200_OK) throw std::runtime_error(std::string("dcma_b200: ") + err);

        auto def_coll = marshal_volume_to_collection(def, stationary);
        return deformation_field(std::move(def_coll));
    } catch (const std::exception &e) {
        YLOGWARN("Demons registration failed: " << e.what());
    }
    return std::nullopt;
}
