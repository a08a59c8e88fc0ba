// Host logic of the drop-in AlignViaDemons (dicomautomaton_b200/host/Alignment_Demons.cc) against the stand-in library
// tests/native/fake_register_capi.cc: what the C call is given (grids, volumes) and what becomes of the field it returns.
//   1. regular stacks, no histogram matching -> ONE call (dcma_b200_demons_register_stacks); the field arrives in chunks
//      of 1000 voxels (slices are 24 x 20 = 480 voxels: chunks straddle slices) and every voxel of the deformation_field
//      must hold the value the library wrote for it;
//   2. histogram matching on -> the two-call route, same check;
//   3. a moving stack on another grid reaches the library on ITS grid (resampling is the library's job there).
// Exit 0 = all checks passed.
#include <cmath>
#include <cstdio>

#include "Alignment_Demons.h"
#include "dcma_b200/demons_c_api.h"

extern "C" {
extern dcma_b200_grid g_fake_fixed_grid, g_fake_moving_grid;
extern double g_fake_fixed_sum, g_fake_moving_sum;
extern int g_fake_calls_stacks, g_fake_calls_plain;
extern int64_t g_fake_chunk;
double fake_field_value(int64_t voxel, int c);
}

static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { std::printf("CHECK failed line %d: %s\n", __LINE__, #c); ++g_fail; } } while (0)

static planar_image_collection<float, double> stack(int64_t S, int64_t R, int64_t C, double dx, double dy, double dz, const vec3<double> &o,
                                                    double scale) {
    planar_image_collection<float, double> coll;
    for (int64_t s = 0; s < S; ++s) {
        planar_image<float, double> img;
        img.init_orientation(vec3<double>(1, 0, 0), vec3<double>(0, 1, 0));
        img.init_buffer(R, C, 1);
        img.init_spatial(dx, dy, dz, vec3<double>(0, 0, 0), o + vec3<double>(0, 0, 1) * (dz * static_cast<double>(s)));
        for (int64_t r = 0; r < R; ++r)
            for (int64_t c = 0; c < C; ++c) img.reference(r, c, 0) = static_cast<float>(scale * ((s * R + r) * C + c));
        coll.images.push_back(img);
    }
    return coll;
}

static void check_field(const deformation_field &f, int64_t S, int64_t R, int64_t C) {
    const auto &imgs = f.get_imagecoll_crefw().get().images;
    CHECK(static_cast<int64_t>(imgs.size()) == S);
    int64_t z = 0, bad = 0;
    for (const auto &img : imgs) {
        CHECK(img.rows == R && img.columns == C && img.channels == 3);
        for (int64_t r = 0; r < R; ++r)
            for (int64_t c = 0; c < C; ++c)
                for (int ch = 0; ch < 3; ++ch)
                    if (img.value(r, c, ch) != fake_field_value((z * R + r) * C + c, ch)) ++bad;
        ++z;
    }
    CHECK(bad == 0);
}

int main() {
    const int64_t S = 11, R = 24, C = 20;
    auto fixed = stack(S, R, C, 1.0, 1.0, 2.0, vec3<double>(3, 4, 5), 1.0);
    auto moving = stack(S, R, C, 1.0, 1.0, 2.0, vec3<double>(3, 4, 5), 0.5);
    AlignViaDemonsParams p;
    p.verbosity = 0;
    p.max_iterations = 3;
    {   // 1. one call, chunks straddling slices; also a chunk size of one voxel and one larger than the volume
        for (int64_t chunk : {int64_t(1000), int64_t(1), int64_t(1) << 40, int64_t(480), int64_t(479)}) {
            g_fake_chunk = chunk;
            const int before = g_fake_calls_stacks;
            auto res = AlignViaDemons(p, moving, fixed);
            CHECK(res.has_value());
            CHECK(g_fake_calls_stacks == before + 1);
            if (res) check_field(*res, S, R, C);
        }
        g_fake_chunk = 1000;
        const double n = static_cast<double>(S * R * C);
        CHECK(std::fabs(g_fake_fixed_sum - n * (n - 1) / 2) < 1e-3 * n && std::fabs(g_fake_moving_sum - 0.5 * n * (n - 1) / 2) < 1e-3 * n);
        CHECK(g_fake_fixed_grid.geom.slices == S && g_fake_fixed_grid.geom.rows == R && g_fake_fixed_grid.geom.cols == C);
        CHECK(g_fake_fixed_grid.geom.pxl_dz == 2.0 && g_fake_fixed_grid.origin[0] == 3.0 && g_fake_fixed_grid.origin[2] == 5.0);
    }
    {   // 2. histogram matching needs the resampled image on the host: resample + match + the plain call, then one fill
        p.use_histogram_matching = true;
        const int s0 = g_fake_calls_stacks, p0 = g_fake_calls_plain;
        auto res = AlignViaDemons(p, moving, fixed);
        CHECK(res.has_value() && g_fake_calls_stacks == s0 && g_fake_calls_plain == p0 + 1);
        if (res) check_field(*res, S, R, C);
        p.use_histogram_matching = false;
    }
    {   // 3. a moving stack on ANOTHER grid travels on its own grid
        auto other = stack(7, 15, 13, 1.5, 1.25, 3.0, vec3<double>(-1, 2, 4), 1.0);
        auto res = AlignViaDemons(p, other, fixed);
        CHECK(res.has_value());
        if (res) check_field(*res, S, R, C);
        CHECK(g_fake_moving_grid.geom.slices == 7 && g_fake_moving_grid.geom.rows == 15 && g_fake_moving_grid.geom.cols == 13);
        CHECK(g_fake_moving_grid.geom.pxl_dx == 1.5 && g_fake_moving_grid.geom.pxl_dy == 1.25 && g_fake_moving_grid.geom.pxl_dz == 3.0);
        CHECK(g_fake_moving_grid.origin[0] == -1.0 && g_fake_moving_grid.origin[1] == 2.0 && g_fake_moving_grid.origin[2] == 4.0);
    }
    {   // 4. slice thickness != slice spacing: the engine must see the FIRST image's pxl_dz (Alignment_Demons.h:160-162), so
        //    the single call (one grid for both purposes) is not used
        auto thick = stack(S, R, C, 1.0, 1.0, 2.0, vec3<double>(3, 4, 5), 1.0);
        for (auto &img : thick.images) img.pxl_dz = 2.5;
        const int s0 = g_fake_calls_stacks, p0 = g_fake_calls_plain;
        auto res = AlignViaDemons(p, moving, thick);
        (void)res;
        CHECK(g_fake_calls_stacks == s0 && (g_fake_calls_plain == p0 + 1 || !res.has_value()));
    }
    std::printf(g_fail ? "FAILED %d checks\n" : "ALL OK\n", g_fail);
    return g_fail ? 1 : 0;
}
