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
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <exception>
#include <future>
#include <thread>
#include <vector>

#include "dcma_b200/demons_c_api.h"

#include "Dense_Stack.h"

using namespace AlignViaDemonsHelpers;

using namespace dcma_b200_host::detail;

planar_image_collection<float, double> AlignViaDemonsHelpers::resample_image_to_reference_grid(
    const planar_image_collection<float, double> &moving, const planar_image_collection<float, double> &reference) {
    if (moving.images.empty() || reference.images.empty())
        throw std::invalid_argument("Cannot resample: image collection is empty");
    const float nan = std::numeric_limits<float>::quiet_NaN();
    // Regular stacks (the only kind AlignViaDemons accepts, Alignment_Demons.h:150-152): the per-voxel lookups of
    // cc:722-751 run on the GPU (dcma_b200_resample_to_grid), not as a serial host loop over every fixed-grid voxel.
    {
        dcma_b200_grid mg, rg;
        if (grid_of(moving, mg) && grid_of(reference, rg) && mg.geom.channels == rg.geom.channels) {
            const size_t nm = static_cast<size_t>(mg.geom.slices * mg.geom.rows * mg.geom.cols * mg.geom.channels);
            const size_t nr = static_cast<size_t>(rg.geom.slices * rg.geom.rows * rg.geom.cols * mg.geom.channels);
            Pinned<float> src(nm), dst(nr);
            marshal_into(moving, src.data());
            char err[512] = {0};
            check_rc(dcma_b200_resample_to_grid(&mg, src.data(), &rg, nan, dst.data(), -1, err, sizeof err), err);
            return collection_from_dense<float>(dst.data(), rg.geom.slices, rg.geom.rows, rg.geom.cols, mg.geom.channels, reference);
        }
    }
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
    // Regular stacks: the per-voxel body (cc:921-936: field.transform(pos), then a trilinear lookup of the image at the
    // displaced position, NaN outside) runs on the GPU -- dcma_b200_warp_image_grid, the same kernel WarpImages binds.
    {
        dcma_b200_grid ig, fg;
        const auto &fcoll = def_field.get_imagecoll_crefw().get();
        if (grid_of(img_coll, ig) && grid_of(fcoll, fg) && fg.geom.channels == 3) {
            const size_t ni = static_cast<size_t>(ig.geom.slices * ig.geom.rows * ig.geom.cols * ig.geom.channels);
            const size_t nf = static_cast<size_t>(fg.geom.slices * fg.geom.rows * fg.geom.cols * 3);
            Pinned<float> src(ni), dst(ni);
            Pinned<double> fld(nf);
            marshal_into(img_coll, src.data());
            marshal_into(fcoll, fld.data());
            char err[512] = {0};
            check_rc(dcma_b200_warp_image_grid(&fg, fld.data(), &ig, src.data(), nullptr, std::numeric_limits<float>::quiet_NaN(),
                                               dst.data(), -1, err, sizeof err), err);
            return collection_from_dense<float>(dst.data(), ig.geom.slices, ig.geom.rows, ig.geom.cols, ig.geom.channels, img_coll);
        }
    }
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
        // ---- dense volumes in pinned host memory.  Regular stacks: the moving image is resampled onto the stationary
        // grid on the GPU straight from its own dense volume (no intermediate image collection); anything else takes
        // the reference's route (resample_image_to_reference_grid, then marshal) and fails in marshal exactly as the
        // reference does when the result is not a regular grid.
        if (params.verbosity >= 1) YLOGINFO("Resampling moving image to reference grid");
        // DCMA_B200_TRACE: host-side phase times of this call on stderr (diagnostics; the library prints its own line)
        const bool trace = std::getenv("DCMA_B200_TRACE") != nullptr;
        using clk = std::chrono::steady_clock;
        auto t_prev = clk::now();
        auto lap = [&](const char *what) {
            if (!trace) return;
            const auto now = clk::now();
            std::fprintf(stderr, "[AlignViaDemons] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
            t_prev = now;
        };
        dcma_b200_grid fg, mg;
        demons_volume<float> fixed_vol;  // header only (extents, spacing); the voxels live in the pinned buffers
        Pinned<float> fixed_px, moving_px;
        auto header_from = [](demons_volume<float> &v, const dcma_b200_grid &gr, int64_t channels) {
            v.slices = gr.geom.slices;
            v.rows = gr.geom.rows;
            v.cols = gr.geom.cols;
            v.channels = channels;
            v.pxl_dx = gr.geom.pxl_dx;
            v.pxl_dy = gr.geom.pxl_dy;
            v.pxl_dz = gr.geom.pxl_dz;
        };
        const bool fixed_regular = grid_of(stationary, fg);
        size_t nfix = 0;
        if (fixed_regular) {
            header_from(fixed_vol, fg, fg.geom.channels);
            // the reference takes the spacing of the FIRST image (Alignment_Demons.h:160-162)
            fixed_vol.pxl_dz = stationary.images.front().pxl_dz;
            nfix = static_cast<size_t>(fg.geom.slices * fg.geom.rows * fg.geom.cols * fg.geom.channels);
            fixed_px = Pinned<float>(nfix);
            marshal_into(stationary, fixed_px.data());
            lap("marshal fixed (pinned)");
        } else {
            // not a stack the GPU resampler can describe (e.g. slices out of order in the list): the reference's own
            // marshalling decides whether it is acceptable at all, and throws its diagnostics if not
            auto fv = marshal_collection_to_volume(stationary);
            nfix = fv.data.size();
            fixed_px = Pinned<float>(nfix);
            std::memcpy(fixed_px.data(), fv.data.data(), nfix * sizeof(float));
            fv.data.clear();
            fixed_vol = fv;
        }
        // Both stacks regular and no histogram matching (which needs the resampled image on the host): ONE call resamples on
        // the device, iterates and streams the field back in chunks (dcma_b200_demons_register_stacks, below).
        const bool moving_regular = fixed_regular && grid_of(moving_in, mg) && mg.geom.channels == fg.geom.channels;
        // (the engine takes its slice spacing from the FIRST image's pxl_dz, Alignment_Demons.h:160-162, the resampling from the
        // slice POSITIONS: the single call describes both with one grid, so it is used when the two agree)
        const bool one_call = moving_regular && !params.use_histogram_matching && fixed_vol.pxl_dz == fg.geom.pxl_dz;
        Pinned<float> raw;
        if (moving_regular) {
            const size_t nmov = static_cast<size_t>(mg.geom.slices * mg.geom.rows * mg.geom.cols * mg.geom.channels);
            raw = Pinned<float>(nmov);
            marshal_into(moving_in, raw.data());
            lap("marshal moving (pinned)");
        }
        if (!one_call) moving_px = Pinned<float>(nfix);
        if (one_call) {
            // resampled on the device inside the registration call
        } else if (moving_regular) {
            char rerr[512] = {0};
            check_rc(dcma_b200_resample_to_grid(&mg, raw.data(), &fg, std::numeric_limits<float>::quiet_NaN(), moving_px.data(),
                                                params.device, rerr, sizeof rerr), rerr);
            lap("resample moving (GPU)");
        } else {
            auto moving = resample_image_to_reference_grid(moving_in, stationary);
            const auto mv = marshal_collection_to_volume(moving);
            if (mv.data.size() != nfix) throw std::invalid_argument("Resampled moving image does not match the stationary grid");
            std::memcpy(moving_px.data(), mv.data.data(), nfix * sizeof(float));
        }
        if (params.use_histogram_matching) {
            // reference: histogram_match(moving, stationary, ...) on the collections (cc:967-972).  Here the same
            // algorithm runs on the GPU over the dense volumes (bit-identical result; the public helper above stays a
            // host function, as in the reference)
            if (params.verbosity >= 1) YLOGINFO("Applying histogram matching");
            Pinned<float> matched(nfix);
            char herr[512] = {0};
            int32_t mapped = 0;
            check_rc(dcma_b200_histogram_match(moving_px.data(), static_cast<int64_t>(nfix), fixed_px.data(), static_cast<int64_t>(nfix),
                                               params.histogram_bins, params.histogram_outlier_fraction, matched.data(), &mapped,
                                               params.device, herr, sizeof herr), herr);
            if (!mapped) YLOGWARN("Histogram matching not applicable (no finite values or constant intensity), moving image unchanged");
            moving_px = std::move(matched);
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
        if (const char *e = std::getenv("DCMA_B200_FIELD_MODE")) {  // test matrix / site policy override: never silent
            if (!std::strcmp(e, "fp64")) p.field_mode = DCMA_B200_FIELD_FP64;
            else if (!std::strcmp(e, "fp64_strict")) p.field_mode = DCMA_B200_FIELD_FP64_STRICT;
            else if (!std::strcmp(e, "fp32")) p.field_mode = DCMA_B200_FIELD_FP32;
            else if (!std::strcmp(e, "mixed")) p.field_mode = DCMA_B200_FIELD_MIXED;
            else throw std::invalid_argument(std::string("DCMA_B200_FIELD_MODE='") + e + "' is not one of fp64, fp64_strict, fp32, mixed");
            YLOGWARN("DCMA_B200_FIELD_MODE=" << e << " overrides the caller's field_mode");
        }
        p.device = params.device;
        p.pyramid_levels = params.pyramid_levels;
        p.use_symmetric_force = params.use_symmetric_force ? 1 : 0;
        for (int i = 0; i < 3; ++i) p.pyramid_iterations[i] = params.pyramid_iterations[i];
        p.n_devices = params.n_devices;
        if (const char *e = std::getenv("DCMA_B200_DEVICES")) {  // site policy: GPUs to shard over; never silent
            char *end = nullptr;
            const long n = std::strtol(e, &end, 10);
            if (end == e || *end != '\0' || n < 1 || n > 64)
                throw std::invalid_argument(std::string("DCMA_B200_DEVICES='") + e + "' is not a device count");
            p.n_devices = static_cast<int32_t>(n);
            YLOGWARN("DCMA_B200_DEVICES=" << e << " overrides the caller's n_devices");
        }
        p.halo_planes = params.halo_planes;

        const size_t nvox = static_cast<size_t>(fixed_vol.slices * fixed_vol.rows * fixed_vol.cols);
        Pinned<double> def(nvox * 3);
        char err[512] = {0};
        dcma_b200_run_stats stats;
        // marshal_volume_to_collection (Alignment_Demons.h:182-204) in two steps: the output images are built (allocated,
        // first-touched) by a helper thread WHILE the GPU iterates; the copy runs when the field has arrived
        planar_image_collection<double, double> def_coll;
        std::exception_ptr alloc_error;
        std::thread builder([&]() {
            try {
                def_coll = collection_like<double>(fixed_vol.slices, fixed_vol.rows, fixed_vol.cols, 3, stationary);
            } catch (...) {
                alloc_error = std::current_exception();
            }
        });
        // Chunks of the field are copied into the output stack as they land, by helper tasks, while the rest of the field
        // is still crossing PCIe.  The callback runs on this thread inside the C call: it must not throw.
        struct ChunkCopy {
            std::thread *builder;
            planar_image_collection<double, double> *coll;
            const double *src;
            int64_t R, C;
            std::vector<planar_image<double, double> *> imgs;
            std::vector<std::future<void>> tasks;
            std::exception_ptr error;
            bool dense = false;
            void ready() {
                if (builder->joinable()) builder->join();
                if (!imgs.empty()) return;
                for (auto &img : coll->images) imgs.push_back(&img);
                if (imgs.empty()) return;
                const auto &a = *imgs.front();
                dense = (a.index(0, 0, 0) == 0) && (a.index(0, 0, 1) == 1) && (C == 1 || a.index(0, 1, 0) == 3) &&
                        (R == 1 || a.index(1, 0, 0) == C * 3) && a.data.size() == static_cast<size_t>(R * C * 3);
            }
            void copy(int64_t v0, int64_t v1) {  // voxels [v0, v1) of the dense field -> the images they belong to
                const int64_t per = R * C;
                for (int64_t v = v0; v < v1;) {
                    const int64_t z = v / per, in0 = v - z * per, n = std::min(per - in0, v1 - v);
                    auto &img = *imgs.at(static_cast<size_t>(z));
                    const double *s = src + 3 * v;
                    if (dense) {
                        std::memcpy(img.data.data() + 3 * in0, s, static_cast<size_t>(3 * n) * sizeof(double));
                    } else {
                        for (int64_t i = 0; i < n; ++i) {
                            const int64_t y = (in0 + i) / C, x = (in0 + i) - y * C;
                            for (int64_t c = 0; c < 3; ++c) img.reference(y, x, c) = s[3 * i + c];
                        }
                    }
                    v += n;
                }
            }
            static void on_chunk(void *user, int64_t v0, int64_t v1) {
                auto *self = static_cast<ChunkCopy *>(user);
                try {
                    self->ready();
                    if (self->imgs.empty()) return;  // the stack could not be built: reported after the call
                    // a 48 MiB chunk in 4 pieces: one thread copies at a fraction of what the PCIe link delivers
                    const int64_t step = std::max<int64_t>((v1 - v0 + 3) / 4, 1);
                    for (int64_t a = v0; a < v1; a += step) {
                        const int64_t b = std::min(v1, a + step);
                        self->tasks.push_back(std::async(std::launch::async, [self, a, b]() { self->copy(a, b); }));
                    }
                } catch (...) {
                    if (!self->error) self->error = std::current_exception();
                }
            }
        } chunks{&builder, &def_coll, def.data(), fixed_vol.rows, fixed_vol.cols, {}, {}, nullptr, false};
        int rc = DCMA_B200_OK;
        if (one_call) {
            rc = dcma_b200_demons_register_stacks(&fg, fixed_px.data(), &mg, raw.data(), &p, def.data(), nullptr, &stats,
                                                  &ChunkCopy::on_chunk, &chunks, err, sizeof(err));
        } else {
            rc = dcma_b200_demons_register(&g, fixed_px.data(), moving_px.data(), &p, def.data(), nullptr, &stats, err, sizeof(err));
        }
        lap(one_call ? "dcma_b200_demons_register_stacks" : "dcma_b200_demons_register");
        if (builder.joinable()) builder.join();
        for (auto &t : chunks.tasks) t.get();
        lap("wait for the output stack");
        check_rc(rc, err);
        if (alloc_error) std::rethrow_exception(alloc_error);
        if (chunks.error) std::rethrow_exception(chunks.error);
        if (!one_call) fill_from_dense(def_coll, def.data(), fixed_vol.rows, fixed_vol.cols, 3);
        lap("copy field into the stack");
        deformation_field out(std::move(def_coll));
        lap("deformation_field(...)");
        return out;
    } catch (const std::exception &e) {
        YLOGWARN("Demons registration failed: " << e.what());
    }
    return std::nullopt;
}
