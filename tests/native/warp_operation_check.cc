// Drives the reference's OWN operation file src/Operations/WarpImages.cc -- compiled in place by
// dicomautomaton_b200/host/build_host_tests.py::build_warp_operation, unmodified or with patches/WarpImages.patch applied --
// the way the dispatcher does: arguments as strings, defaults from OpArgDocWarpImages().
//   warp_operation_check doc   prints "name=default" for every documented argument
//   warp_operation_check run   warps a dose-like stack (ImageSelection, its own grid) onto the grid of a second stack
//                              (ReferenceImageSelection) with a deformation_field estimated on a third grid, inside one
//                              rectangular ROI per slice, and checks what the operation leaves in the Drover
//                              (WarpImages.cc:385-468): a new image array with the reference geometry; voxels outside
//                              the ROI keep the placeholder's values; voxels inside hold source(p + D(p)).
// The expected values are computed here with the stand-in's deformation_field::transform and adjacency look-up, i.e. the
// statements of WarpImages.cc:424-452.  Exit 0 = all checks passed, 3 = the operation threw (message printed; the patched
// build without a device: "dcma_b200: no CUDA device ...", there is no CPU fallback for an accepted request).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>

#include "Structs.h"
#include "Regex_Selectors.h"

OperationDoc OpArgDocWarpImages();
bool WarpImages(Drover &, const OperationArgPkg &, std::map<std::string, std::string> &, const std::string &);

static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { std::printf("CHECK failed line %d: %s\n", __LINE__, #c); ++g_fail; } } while (0)

template <class T, class F>
static planar_image_collection<T, double> make_stack(int64_t S, int64_t R, int64_t C, int64_t CH, double dx, double dy, double dz,
                                                     const vec3<double> &origin, F value) {
    planar_image_collection<T, double> coll;
    for (int64_t s = 0; s < S; ++s) {
        planar_image<T, double> img;
        img.init_orientation(vec3<double>(1, 0, 0), vec3<double>(0, 1, 0));
        img.init_buffer(R, C, CH);
        img.init_spatial(dx, dy, dz, vec3<double>(0, 0, 0), origin + vec3<double>(0, 0, 1) * (dz * static_cast<double>(s)));
        for (int64_t r = 0; r < R; ++r)
            for (int64_t c = 0; c < C; ++c)
                for (int64_t ch = 0; ch < CH; ++ch) img.reference(r, c, ch) = static_cast<T>(value(img.position(r, c), ch));
        img.metadata["FrameOfReferenceUID"] = "1.2.3.4";
        coll.images.push_back(img);
    }
    return coll;
}

int main(int argc, char **argv) {
    const OperationDoc doc = OpArgDocWarpImages();
    if (argc > 1 && !std::strcmp(argv[1], "doc")) {
        std::printf("operation=%s\n", doc.name.c_str());
        for (const auto &a : doc.args) std::printf("%s=%s\n", a.name.c_str(), a.default_val.c_str());
        return 0;
    }
    Drover d;
    // #0 the image to warp: a smooth "dose" on a 2 x 2 x 2.5 mm grid
    auto src = std::make_shared<Image_Array>();
    src->imagecoll = make_stack<float>(14, 30, 32, 1, 2.0, 2.0, 2.5, vec3<double>(-4.0, -3.0, -2.0), [](const vec3<double> &p, int64_t) {
        return 60.0 * std::exp(-((p.x - 28.0) * (p.x - 28.0) + (p.y - 26.0) * (p.y - 26.0) + (p.z - 14.0) * (p.z - 14.0)) / 260.0) + 0.2 * p.x + 0.1 * p.y;
    });
    d.image_data.push_back(src);
    // #1 the geometry to warp onto: 1.5 x 1.5 x 3 mm, strictly inside both other grids; its own voxels are -3 (recognisable)
    auto ref = std::make_shared<Image_Array>();
    ref->imagecoll = make_stack<float>(8, 24, 26, 1, 1.5, 1.5, 3.0, vec3<double>(8.0, 7.0, 4.0), [](const vec3<double> &, int64_t) { return -3.0; });
    d.image_data.push_back(ref);
    // the transform: a field on a 1 mm grid covering the reference stack, up to ~2.5 mm of displacement
    auto fld = make_stack<double>(34, 48, 52, 3, 1.0, 1.0, 1.0, vec3<double>(2.0, 1.0, 0.0), [](const vec3<double> &p, int64_t ch) {
        const double a = std::sin(0.11 * p.x + 0.3) * std::cos(0.09 * p.y), b = std::cos(0.07 * p.z + 0.2 * ch);
        return (ch == 0 ? 2.0 : ch == 1 ? -1.5 : 1.0) * a * b;
    });
    auto t3 = std::make_shared<Transform3>();
    t3->transform = deformation_field(std::move(fld));
    d.trans_data.push_back(t3);
    // one ROI: a rectangle on every reference slice (centre rule: voxels whose centre is inside are bounded)
    d.contour_data = std::make_shared<Contour_Data>();
    d.contour_data->ccs.emplace_back();
    const double x0 = 13.1, x1 = 38.2, y0 = 11.3, y1 = 35.7;
    for (const auto &img : ref->imagecoll.images) {
        contour_of_points<double> c;
        const double z = img.position(0, 0).z;
        c.points = {vec3<double>(x0, y0, z), vec3<double>(x1, y0, z), vec3<double>(x1, y1, z), vec3<double>(x0, y1, z)};
        c.metadata["ROIName"] = "target";
        c.metadata["NormalizedROIName"] = "target";
        d.contour_data->ccs.back().contours.push_back(c);
    }

    OperationArgPkg args("WarpImages");
    for (const auto &a : doc.args) args.insert(a.name, a.default_val);
    args.set("ImageSelection", "#0");
    args.set("ReferenceImageSelection", "#1");
    args.set("TransformSelection", "last");
    args.set("ROILabelRegex", "target");
    std::map<std::string, std::string> meta;
    try {
        if (!WarpImages(d, args, meta, "")) { std::printf("operation returned false\n"); return 2; }
    } catch (const std::exception &e) {
        std::printf("THROW: %s\n", e.what());
        return 3;
    }
    CHECK(d.image_data.size() == 3);
    const auto &out = d.image_data.back()->imagecoll;
    CHECK(out.images.size() == ref->imagecoll.images.size());
    const auto &field = std::get<deformation_field>(t3->transform);
    std::list<std::reference_wrapper<planar_image<float, double>>> none;
    planar_image_adjacency<float, double> adj(none, {std::ref(src->imagecoll)}, src->imagecoll.images.front().ortho_unit());
    long inside = 0, outside = 0, moved = 0;
    double worst = 0.0;
    auto io = out.images.begin();
    auto ir = ref->imagecoll.images.begin();
    for (; io != out.images.end() && ir != ref->imagecoll.images.end(); ++io, ++ir) {
        CHECK(io->rows == ir->rows && io->columns == ir->columns && io->channels == ir->channels);
        CHECK(io->position(0, 0) == ir->position(0, 0));
        for (int64_t r = 0; r < io->rows; ++r)
            for (int64_t c = 0; c < io->columns; ++c) {
                const auto p = io->position(r, c);
                const bool in = (p.x > x0 && p.x < x1 && p.y > y0 && p.y < y1);
                const float got = io->value(r, c, 0);
                if (!in) {
                    ++outside;
                    if (got != -3.0f) { CHECK(got == -3.0f); r = io->rows; break; }
                    continue;
                }
                ++inside;
                const float want = adj.trilinearly_interpolate(field.transform(p), 0, 0.0f);   // WarpImages.cc:442-452
                worst = std::max(worst, static_cast<double>(std::fabs(got - want)));
                if (std::fabs(want - adj.trilinearly_interpolate(p, 0, 0.0f)) > 0.05) ++moved;
            }
    }
    CHECK(inside > 1000 && outside > 1000 && moved > inside / 4);
    CHECK(worst <= 2e-4);
    std::printf("bounded voxels %ld, unbounded %ld, displaced noticeably %ld, max |got - expected| = %.3e\n", inside, outside, moved, worst);
    std::printf(g_fail ? "FAILED %d checks\n" : "ALL OK\n", g_fail);
    return g_fail ? 1 : 0;
}
