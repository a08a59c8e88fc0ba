// The reference's own cases for the field text format (src/Alignment_Field_Tests.cc:196-289), run against the
// reference's own write_to / read_from with patches/Alignment_Field.patch applied (numbers converted on the GPU by
// host/Field_Payload.cc; built by host/build_host_tests.py::build_field_io), plus a byte-for-byte comparison with
// the unpatched statement sequence executed on the host with iostreams (src/Alignment_Field.cc:322-357).
// Exit code 0 = all passed, 77 = no device.
#include <cmath>
#include <cstdio>
#include <functional>
#include <iostream>
#include <limits>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "YgorBase64.h"
#include "YgorImages.h"
#include "YgorMath.h"

#include "Alignment_Field.h"
#include "Field_Payload.h"
#include "dcma_b200/demons_c_api.h"

static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { std::printf("CHECK failed line %d: %s\n", __LINE__, #c); ++g_fail; } } while (0)
static bool approx(double a, double b) { return std::fabs(a - b) <= 1e-12 * (1.0 + std::fabs(a) + std::fabs(b)); }

// mirrors make_field_test_vector_field (src/Alignment_Field_Tests.cc:43-80)
static planar_image_collection<double, double> make_field(int64_t slices, int64_t rows, int64_t cols,
                                                          const std::function<vec3<double>(int64_t, int64_t, int64_t)> &fn) {
    planar_image_collection<double, double> coll;
    for (int64_t s = 0; s < slices; ++s) {
        planar_image<double, double> img;
        img.init_orientation(vec3<double>(1, 0, 0), vec3<double>(0, 1, 0));
        img.init_buffer(rows, cols, 3);
        img.init_spatial(1.0, 1.0, 1.0, vec3<double>(0, 0, 0), vec3<double>(0, 0, 1) * static_cast<double>(s));
        for (int64_t r = 0; r < rows; ++r)
            for (int64_t c = 0; c < cols; ++c) {
                const auto v = fn(s, r, c);
                img.reference(r, c, 0) = v.x; img.reference(r, c, 1) = v.y; img.reference(r, c, 2) = v.z;
            }
        coll.images.push_back(img);
    }
    return coll;
}

// the reference's write_to, statement by statement, with host iostreams
static std::string host_write(const planar_image_collection<double, double> &field) {
    std::ostringstream os;
    os.precision(std::numeric_limits<double>::max_digits10);
    os << static_cast<int64_t>(field.images.size()) << '\n';
    for (const auto &img : field.images) {
        os << img.rows << " " << img.columns << " " << img.channels << '\n';
        os << img.pxl_dx << " " << img.pxl_dy << " " << img.pxl_dz << '\n';
        os << img.anchor << '\n' << img.offset << '\n' << img.row_unit << '\n' << img.col_unit << '\n';
        os << "num_metadata= " << img.metadata.size() << '\n';
        for (const auto &mp : img.metadata) os << Base64::EncodeFromString(mp.first) << " " << Base64::EncodeFromString(mp.second) << '\n';
        for (const auto &val : img.data) os << val << '\n';
    }
    return os.str();
}

int main() {
    if (dcma_b200_device_count() < 1) { std::printf("SKIP no device\n"); return 77; }
    {   // "deformation_field write_to and read_from roundtrip" (:196-248)
        auto imgs = make_field(2, 3, 4, [](int64_t s, int64_t r, int64_t c) { return vec3<double>(0.1 * c, -0.2 * r, 0.05 * s); });
        for (auto &img : imgs.images) { img.metadata["PatientID"] = "test_patient_001"; img.metadata["key with spaces"] = "value;with@special=chars"; }
        const std::string want = host_write(imgs);
        deformation_field field(std::move(imgs));
        std::stringstream ss;
        CHECK(field.write_to(ss));
        CHECK(ss.str() == want);
        deformation_field field2(ss);
        const auto &a = field.get_imagecoll_crefw().get().images, &b = field2.get_imagecoll_crefw().get().images;
        CHECK(a.size() == b.size());
        auto ia = a.begin(); auto ib = b.begin();
        for (; ia != a.end() && ib != b.end(); ++ia, ++ib) {
            CHECK(ia->rows == ib->rows); CHECK(ia->columns == ib->columns); CHECK(ia->channels == ib->channels);
            CHECK(approx(ia->pxl_dx, ib->pxl_dx)); CHECK(approx(ia->pxl_dy, ib->pxl_dy)); CHECK(approx(ia->pxl_dz, ib->pxl_dz));
            CHECK(ia->data == ib->data);  // stronger than the reference's Approx: bit-identical
            CHECK(ia->metadata == ib->metadata);
        }
        const vec3<double> p(1.5, 0.5, 0.5);
        const auto t1 = field.transform(p), t2 = field2.transform(p);
        CHECK(approx(t1.x, t2.x)); CHECK(approx(t1.y, t2.y)); CHECK(approx(t1.z, t2.z));
    }
    {   // "preserves zero field" (:251-265)
        deformation_field field(make_field(1, 2, 2, [](int64_t, int64_t, int64_t) { return vec3<double>(0, 0, 0); }));
        std::stringstream ss;
        CHECK(field.write_to(ss));
        deformation_field field2(ss);
        const vec3<double> p(0.5, 0.5, 0.0);
        const auto r = field2.transform(p);
        CHECK(approx(r.x, p.x)); CHECK(approx(r.y, p.y)); CHECK(approx(r.z, p.z));
    }
    {   // "read_from rejects invalid input" (:268-289)
        for (const char *bad : {"", "-1\n", "0\n", "1\n3 3 3\n"}) {
            std::stringstream ss(bad);
            bool threw = false;
            try { deformation_field f(ss); } catch (const std::exception &) { threw = true; }
            CHECK(threw);
        }
        // pixel data that ends early / is not a number
        auto imgs = make_field(1, 2, 2, [](int64_t, int64_t r, int64_t c) { return vec3<double>(r, c, 0.5); });
        std::string good = host_write(imgs);
        for (std::string bad : {good.substr(0, good.size() - 4), good.substr(0, good.size() - 2) + "x\n"}) {
            std::stringstream ss(bad);
            bool threw = false;
            try { deformation_field f(ss); } catch (const std::exception &) { threw = true; }
            CHECK(threw);
        }
    }
    {   // a larger random field: text identical to the host iostream writer, values identical after reading back
        std::mt19937_64 rng(7);
        std::normal_distribution<double> nd(0.0, 2.5);
        auto imgs = make_field(5, 64, 48, [&](int64_t, int64_t, int64_t) { return vec3<double>(nd(rng), nd(rng) * 1e-7, std::ldexp(static_cast<double>(rng() >> 40), -20)); });
        const std::string want = host_write(imgs);
        planar_image_collection<double, double> copy_of_imgs(imgs);
        deformation_field field(std::move(copy_of_imgs));
        std::stringstream ss;
        CHECK(field.write_to(ss));
        CHECK(ss.str() == want);
        ss << "trailing 42";  // read_from must stop just behind the last number, as n extractions do
        deformation_field back(ss);
        std::string word; int number = 0;
        ss >> word >> number;
        CHECK(word == "trailing"); CHECK(number == 42);
        const auto &bi = back.get_imagecoll_crefw().get().images;
        CHECK(bi.size() == imgs.images.size());
        auto ia = imgs.images.begin(); auto ib = bi.begin();
        for (; ia != imgs.images.end() && ib != bi.end(); ++ia, ++ib) { CHECK(ia->data == ib->data); CHECK(ia->offset == ib->offset); }
    }
    {   // the codec alone: odd white space, long tokens (the bounded read-ahead has to grow), too few numbers
        std::stringstream ss;
        const int n = 300;
        for (int i = 0; i < n; ++i) ss << "\t \n  0000000000000000000000000000000000000" << i << ".2500000000000000000000000000000000000000  \r\n";
        ss << "end";
        std::vector<double> v(n, -1.0);
        CHECK(dcma_b200_io::read_payload(ss, v.data(), n));
        for (int i = 0; i < n; ++i) CHECK(v[i] == i + 0.25);
        std::string word; ss >> word; CHECK(word == "end");
        std::stringstream few("1 2 3");
        CHECK(!dcma_b200_io::read_payload(few, v.data(), 4)); CHECK(few.fail());
        std::stringstream os2;
        const double vals[3] = {0.1, -2.5, 1e300};
        CHECK(dcma_b200_io::write_payload(os2, vals, 3)); CHECK(os2.str() == "0.10000000000000001\n-2.5\n1.0000000000000001e+300\n");
    }
    std::printf(g_fail ? "FAILED %d checks\n" : "ALL OK\n", g_fail);
    return g_fail ? 1 : 0;
}
