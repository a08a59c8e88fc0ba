// TEST stand-in for the C-ABI entry points host/Warp_Images_Field.cc calls (dcma_b200_warp_to_grid and the pinned host
// allocator), computed on the CPU with the definition documented in csrc/warp_grid.cu.  Linked INSTEAD of libdcma_b200.so
// by tests/test_warp_operation.py so that the PATCHED WarpImages operation -- mask order, marshalling of the three stacks,
// channel selection, write-back into the placeholder -- can be exercised on a machine without a GPU.  Never part of the
// product; the GPU test runs the same driver against the real library.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dcma_b200/demons_c_api.h"

namespace {
bool axis(double u, int64_t n, int64_t &i0, int64_t &i1, double &t) {
    if (n == 1) { i0 = i1 = 0; t = 0.0; return std::fabs(u) <= 0.5; }
    if (!(u >= 0.0 && u <= static_cast<double>(n - 1))) return false;
    i0 = std::min<int64_t>(static_cast<int64_t>(std::floor(u)), n - 2);
    i1 = i0 + 1;
    t = u - static_cast<double>(i0);
    return true;
}
void to_index(const dcma_b200_grid &g, const double p[3], double u[3]) {
    const double q[3] = {p[0] - g.origin[0], p[1] - g.origin[1], p[2] - g.origin[2]};
    u[0] = (q[0] * g.ex[0] + q[1] * g.ex[1] + q[2] * g.ex[2]) / g.geom.pxl_dx;
    u[1] = (q[0] * g.ey[0] + q[1] * g.ey[1] + q[2] * g.ey[2]) / g.geom.pxl_dy;
    u[2] = (q[0] * g.ez[0] + q[1] * g.ez[1] + q[2] * g.ez[2]) / g.geom.pxl_dz;
}
template <class V>
double trilinear(const V *d, const dcma_b200_grid &g, int64_t nch, int64_t c, const double u[3], bool &ok) {
    int64_t x0, x1, y0, y1, z0, z1;
    double tx, ty, tz;
    ok = axis(u[0], g.geom.cols, x0, x1, tx) & axis(u[1], g.geom.rows, y0, y1, ty) & axis(u[2], g.geom.slices, z0, z1, tz);
    if (!ok) return 0.0;
    auto at = [&](int64_t z, int64_t y, int64_t x) { return static_cast<double>(d[((z * g.geom.rows + y) * g.geom.cols + x) * nch + c]); };
    const double c00 = at(z0, y0, x0) * (1 - tx) + at(z0, y0, x1) * tx, c10 = at(z0, y1, x0) * (1 - tx) + at(z0, y1, x1) * tx;
    const double c01 = at(z1, y0, x0) * (1 - tx) + at(z1, y0, x1) * tx, c11 = at(z1, y1, x0) * (1 - tx) + at(z1, y1, x1) * tx;
    const double v = (c00 * (1 - ty) + c10 * ty) * (1 - tz) + (c01 * (1 - ty) + c11 * ty) * tz;
    ok = std::isfinite(v);
    return v;
}
}  // namespace

extern "C" {
int dcma_b200_device_count(void) { return 1; }
int dcma_b200_host_alloc(void **out, size_t bytes, char *, size_t) { *out = std::malloc(bytes ? bytes : 1); return *out ? 0 : -3; }
void dcma_b200_host_free(void *p) { std::free(p); }
int dcma_b200_warp_to_grid(const dcma_b200_grid *fg, const double *D, const dcma_b200_grid *sg, const float *src, const dcma_b200_grid *og,
                           const unsigned char *mask, int32_t channel, float oob, float *out, int, char *, size_t) {
    const int64_t nvox = og->geom.slices * og->geom.rows * og->geom.cols, och = og->geom.channels, sch = sg->geom.channels;
    for (int64_t v = 0; v < nvox; ++v) {
        if (mask && !mask[v]) continue;
        const int64_t x = v % og->geom.cols, r = v / og->geom.cols, y = r % og->geom.rows, z = r / og->geom.rows;
        double p[3], u[3], cp[3], us[3];
        for (int k = 0; k < 3; ++k)
            p[k] = og->origin[k] + x * og->geom.pxl_dx * og->ex[k] + y * og->geom.pxl_dy * og->ey[k] + z * og->geom.pxl_dz * og->ez[k];
        to_index(*fg, p, u);
        bool a, b, c;
        const double dx = trilinear(D, *fg, 3, 0, u, a), dy = trilinear(D, *fg, 3, 1, u, b), dz = trilinear(D, *fg, 3, 2, u, c);
        const bool okd = a && b && c;
        for (int k = 0; k < 3; ++k) cp[k] = p[k] + dx * fg->ex[k] + dy * fg->ey[k] + dz * fg->ez[k];
        to_index(*sg, cp, us);
        for (int64_t ch = 0; ch < och; ++ch) {
            if (channel >= 0 && ch != channel) continue;
            bool ok = false;
            const double val = okd ? trilinear(src, *sg, sch, ch, us, ok) : 0.0;
            out[v * och + ch] = (okd && ok) ? static_cast<float>(val) : oob;
        }
    }
    return 0;
}
}
