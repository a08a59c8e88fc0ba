// Drives the reference's OWN operation file (src/Operations/RegisterImagesDemons.cc, compiled unmodified by
// dicomautomaton_b200/host/build_host_tests.py against the AlignViaDemons drop-in and the stand-in headers) the way the
// dispatcher does: arguments as strings, defaults taken from OpArgDocRegisterImagesDemons().
//   operation_check doc   prints "name=default" for every documented argument
//   operation_check run3  (the operation with patches/RegisterImagesDemons.patch applied) the same through the additive
//                         arguments: PyramidLevels=3, UseSymmetricForce=true, FieldMode=fp64 -- BASELINE config 3's shape of run
//   operation_check run   registers a small synthetic pair through RegisterImagesDemons() and checks what the operation
//                         leaves in the Drover (RegisterImagesDemons.cc:337-381).  Exit 0 = all checks passed,
//                         3 = the operation threw (its message is printed: without a device "Demons registration failed")
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>

#include "Structs.h"
#include "Regex_Selectors.h"

OperationDoc OpArgDocRegisterImagesDemons();
bool RegisterImagesDemons(Drover &, const OperationArgPkg &, std::map<std::string, std::string> &, const std::string &);

static std::shared_ptr<Image_Array> make_array(int64_t S, int64_t R, int64_t C, double shift_x) {
    auto ia = std::make_shared<Image_Array>();
    for (int64_t s = 0; s < S; ++s) {
        planar_image<float, double> img;
        img.init_orientation(vec3<double>(1, 0, 0), vec3<double>(0, 1, 0));
        img.init_buffer(R, C, 1);
        img.init_spatial(1.0, 1.0, 1.0, vec3<double>(0, 0, 0), vec3<double>(0, 0, 1) * static_cast<double>(s));
        for (int64_t r = 0; r < R; ++r)
            for (int64_t c = 0; c < C; ++c) {
                const double x = c - shift_x - 0.5 * C, y = r - 0.5 * R, z = s - 0.5 * S;
                const double v = 100.0 * std::exp(-(x * x + y * y + z * z) / 40.0) + 20.0 * std::sin(0.5 * (c - shift_x)) * std::cos(0.4 * r);
                img.reference(r, c, 0) = static_cast<float>(v);
            }
        ia->imagecoll.images.push_back(img);
    }
    return ia;
}

static double mse(const planar_image_collection<float, double> &a, const planar_image_collection<float, double> &b) {
    double s = 0;
    long n = 0;
    auto ia = a.images.begin();
    auto ib = b.images.begin();
    for (; ia != a.images.end() && ib != b.images.end(); ++ia, ++ib)
        for (size_t i = 0; i < ia->data.size(); ++i)
            if (std::isfinite(ia->data[i]) && std::isfinite(ib->data[i])) {
                const double d = ia->data[i] - ib->data[i];
                s += d * d;
                ++n;
            }
    return n ? s / n : 0.0;
}

int main(int argc, char **argv) {
    const OperationDoc doc = OpArgDocRegisterImagesDemons();
    if (argc > 1 && !std::strcmp(argv[1], "doc")) {
        std::printf("operation=%s\n", doc.name.c_str());
        for (const auto &a : doc.args) std::printf("%s=%s\n", a.name.c_str(), a.default_val.c_str());
        return 0;
    }
    OperationArgPkg args(doc.name);
    for (const auto &a : doc.args)
        if (a.expected) args.insert(a.name, a.default_val);  // what the dispatcher does for expected arguments
    args.set("FixedImageSelection", "first");
    args.set("MovingImageSelection", "last");
    args.set("MaxIterations", "40");
    args.set("ConvergenceThreshold", "0");
    args.set("UseDiffeomorphic", "true");
    args.set("KeepDeformedImages", "true");
    args.set("TransformName", "check");
    args.set("Metadata", "Study@phantom;Operator@test");
    args.set("Verbosity", "0");
    const bool run3 = (argc > 1 && !std::strcmp(argv[1], "run3"));
    int64_t S = 12, R = 24, C = 28;
    if (run3) {  // needs the additive arguments of patches/RegisterImagesDemons.patch
        bool have = false;
        for (const auto &a : doc.args) have = have || (a.name == "PyramidLevels");
        if (!have) {
            std::printf("this binary was built from the unpatched operation: no PyramidLevels argument\n");
            return 4;
        }
        args.set("PyramidLevels", "3");
        args.set("PyramidIterations", "15,25,40");
        args.set("UseSymmetricForce", "true");
        args.set("FieldMode", "fp64");
        args.set("MaxIterations", "40");
        S = 32; R = 48; C = 64;  // every axis >= 8 at the coarsest of three levels
    }

    Drover d;
    d.image_data.push_back(make_array(S, R, C, 0.0));  // fixed
    d.image_data.push_back(make_array(S, R, C, 1.5));  // moving: shifted by 1.5 voxels along x
    std::map<std::string, std::string> invocation;
    bool ok = false;
    try {
        ok = RegisterImagesDemons(d, args, invocation, "");
    } catch (const std::exception &e) {
        std::printf("THROW: %s\n", e.what());
        return 3;
    }
    int fail = 0;
#define CHECK(c) do { if (!(c)) { std::printf("CHECK failed line %d: %s\n", __LINE__, #c); ++fail; } } while (0)
    CHECK(ok);
    CHECK(d.trans_data.size() == 1);
    if (d.trans_data.size() == 1) {
        const auto &t = *d.trans_data.front();
        CHECK(std::holds_alternative<deformation_field>(t.transform));
        CHECK(t.metadata.at("RegistrationMethod") == "Demons");
        CHECK(t.metadata.at("RegistrationName") == "check");
        CHECK(t.metadata.at("UseDiffeomorphic") == "true");
        CHECK(t.metadata.at("UseHistogramMatching") == "false");
        CHECK(t.metadata.at("MaxIterations") == "40");
        CHECK(t.metadata.count("DeformationFieldSmoothingSigma") == 1);
        CHECK(t.metadata.at("Study") == "phantom" && t.metadata.at("Operator") == "test");
        if (std::holds_alternative<deformation_field>(t.transform)) {
            const auto &imgs = std::get<deformation_field>(t.transform).get_imagecoll_crefw().get().images;
            CHECK(static_cast<int64_t>(imgs.size()) == S);
            CHECK(imgs.front().rows == R && imgs.front().columns == C && imgs.front().channels == 3);
        }
    }
    CHECK(d.image_data.size() == 3);  // KeepDeformedImages appends the warped moving images (:366-381)
    if (d.image_data.size() == 3) {
        const auto &fixed = d.image_data.front()->imagecoll;
        const auto &moving = (*std::next(d.image_data.begin()))->imagecoll;
        const auto &warped = d.image_data.back()->imagecoll;
        CHECK(static_cast<int64_t>(warped.images.size()) == S);
        CHECK(warped.images.front().metadata.at("Description") == "Demons-warped moving image");
        const double before = mse(fixed, moving), after = mse(fixed, warped);
        std::printf("MSE before %.4f after %.4f\n", before, after);
        CHECK(after < 0.5 * before);
    }
    std::printf(fail ? "FAILED %d checks\n" : "ALL OK\n", fail);
    return fail ? 1 : 0;
}
