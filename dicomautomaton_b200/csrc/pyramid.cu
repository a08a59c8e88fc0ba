// dcma_b200 -- multi-resolution pyramid.  EXTENSION: the reference has no pyramid
// (src/Alignment_Demons.h:20-61 has no such parameter); `pyramid_levels = 1` (the default) is exactly the reference's
// single-resolution loop.  Definition implemented here (restated on the CPU in oracle/extensions.py, "parity unpinned"):
//   * level l+1 halves every axis whose extent at level l is >= 8 (others are kept): extent' = ceil(extent / 2),
//     spacing' = 2 * spacing.  The level-l image is first low-passed along every halved axis (anti-aliasing: a 2x2x2 block
//     mean alone leaves most of the energy between the new and the old Nyquist frequency, which then aliases into wrong
//     forces on the coarse levels): Gaussian, sigma = 1 voxel of level l, 7 taps, weights renormalised over the in-grid
//     finite taps, double accumulation in tap order, rounded to float after each axis (x, then y, then z).  An image
//     voxel of level l+1 is then the mean of the FINITE voxels of its 2x2x2 (2 or 1 along kept axes) block of the
//     low-passed image, accumulated in double in z,y,x order, NaN if the block has no finite voxel;
//   * levels run coarse -> fine with the same parameters (sigmas are in mm, so radii shrink in voxels); level l starts
//     from the level l+1 result prolonged trilinearly: fine voxel x samples the coarse field at u = x/2 - 1/4 (the coarse
//     voxel centre of fine index 2i + 1/2), clamped to [0, extent' - 1]; displacements are in mm and are not rescaled;
//   * the first force evaluation of a warm-started level warps M through the prolonged field (W := M only for D = 0,
//     cc:183);
//   * iterations: max_iterations at the finest level; `pyramid_iterations[l-1]` (0 = max_iterations) at coarser level l.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#include "common.cuh"
#include "engine.h"
#include "pyramid.h"
#include "slab.h"

namespace dcma {

namespace {

struct Dim3i {
    int S, R, C;
};

__device__ __forceinline__ bool finite_f32(float v) { return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u; }

__global__ void __launch_bounds__(256)
    k_restrict(const float *__restrict__ fine, float *__restrict__ coarse, const Dim3i nf, const Dim3i nc, const int ch,
               const int hz, const int hy, const int hx /* 1 = this axis is halved */) {
    const long long total = static_cast<long long>(nc.S) * nc.R * nc.C * ch;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = static_cast<int>(i % ch);
        long long v = i / ch;
        const int x = static_cast<int>(v % nc.C);
        v /= nc.C;
        const int y = static_cast<int>(v % nc.R);
        const int z = static_cast<int>(v / nc.R);
        double sum = 0.0;
        int n = 0;
        for (int dz = 0; dz <= hz; ++dz) {
            const int zf = hz ? 2 * z + dz : z;
            if (zf >= nf.S) continue;
            for (int dy = 0; dy <= hy; ++dy) {
                const int yf = hy ? 2 * y + dy : y;
                if (yf >= nf.R) continue;
                for (int dx = 0; dx <= hx; ++dx) {
                    const int xf = hx ? 2 * x + dx : x;
                    if (xf >= nf.C) continue;
                    const float val = fine[((static_cast<long long>(zf) * nf.R + yf) * nf.C + xf) * ch + c];
                    if (finite_f32(val)) {
                        sum += static_cast<double>(val);
                        ++n;
                    }
                }
            }
        }
        coarse[i] = (n > 0) ? static_cast<float>(sum / static_cast<double>(n)) : __int_as_float(0x7fc00000);
    }
}

// anti-aliasing low-pass along one axis (see the header comment): 7-tap Gaussian, sigma = 1 voxel
__constant__ double c_blur_w[7] = {0.011108996538242306, 0.1353352832366127, 0.6065306597126334, 1.0,
                                   0.6065306597126334, 0.1353352832366127, 0.011108996538242306};
__global__ void __launch_bounds__(256)
    k_blur_axis(const float *__restrict__ in, float *__restrict__ out, const Dim3i n, const int ch, const int axis /* 0 x, 1 y, 2 z */) {
    const long long total = static_cast<long long>(n.S) * n.R * n.C * ch;
    const long long stride_thr = static_cast<long long>(gridDim.x) * blockDim.x;
    const int extent = axis == 0 ? n.C : (axis == 1 ? n.R : n.S);
    const long long step = (axis == 0 ? 1LL : (axis == 1 ? static_cast<long long>(n.C) : static_cast<long long>(n.R) * n.C)) * ch;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride_thr) {
        long long v = i / ch;
        const int x = static_cast<int>(v % n.C);
        v /= n.C;
        const int y = static_cast<int>(v % n.R);
        const int z = static_cast<int>(v / n.R);
        const int pos = axis == 0 ? x : (axis == 1 ? y : z);
        double sum = 0.0, wsum = 0.0;
        for (int k = -3; k <= 3; ++k) {
            const int q = pos + k;
            if (q < 0 || q >= extent) continue;
            const float val = in[i + k * step];
            if (finite_f32(val)) {
                sum = __dadd_rn(sum, __dmul_rn(c_blur_w[k + 3], static_cast<double>(val)));  // no contraction: the CPU restatement has none
                wsum = __dadd_rn(wsum, c_blur_w[k + 3]);
            }
        }
        out[i] = (wsum > 0.0) ? static_cast<float>(sum / wsum) : __int_as_float(0x7fc00000);
    }
}

__device__ __forceinline__ void prolong_axis(int xf, int halved, int nc, int &i0, int &i1, double &t) {
    double u = halved ? 0.5 * static_cast<double>(xf) - 0.25 : static_cast<double>(xf);
    u = fmin(fmax(u, 0.0), static_cast<double>(nc - 1));
    const double f = floor(u);
    i0 = static_cast<int>(f);
    i1 = min(i0 + 1, nc - 1);
    t = u - f;
}

// AoS double fields [z][y][x][3]
__global__ void __launch_bounds__(256)
    k_prolong(const double *__restrict__ coarse, double *__restrict__ fine, const Dim3i nc, const Dim3i nf, const int hz,
              const int hy, const int hx) {
    const long long total = static_cast<long long>(nf.S) * nf.R * nf.C;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < total; v += stride) {
        const int x = static_cast<int>(v % nf.C);
        const long long r = v / nf.C;
        const int y = static_cast<int>(r % nf.R);
        const int z = static_cast<int>(r / nf.R);
        int x0, x1, y0, y1, z0, z1;
        double tx, ty, tz;
        prolong_axis(x, hx, nc.C, x0, x1, tx);
        prolong_axis(y, hy, nc.R, y0, y1, ty);
        prolong_axis(z, hz, nc.S, z0, z1, tz);
        auto at = [&](int zz, int yy, int xx, int c) {
            return coarse[((static_cast<long long>(zz) * nc.R + yy) * nc.C + xx) * 3 + c];
        };
        for (int c = 0; c < 3; ++c) {
            const double c00 = at(z0, y0, x0, c) * (1.0 - tx) + at(z0, y0, x1, c) * tx;
            const double c10 = at(z0, y1, x0, c) * (1.0 - tx) + at(z0, y1, x1, c) * tx;
            const double c01 = at(z1, y0, x0, c) * (1.0 - tx) + at(z1, y0, x1, c) * tx;
            const double c11 = at(z1, y1, x0, c) * (1.0 - tx) + at(z1, y1, x1, c) * tx;
            const double c0 = c00 * (1.0 - ty) + c10 * ty;
            const double c1 = c01 * (1.0 - ty) + c11 * ty;
            fine[v * 3 + c] = c0 * (1.0 - tz) + c1 * tz;
        }
    }
}

int grid_for(long long n) { return static_cast<int>(std::min<long long>((n + 255) / 256, 148LL * 16)); }

template <typename X>
struct DevBuf {
    X *p = nullptr;
    explicit DevBuf(long long n) { DCMA_CUDA(cudaMalloc(&p, sizeof(X) * static_cast<size_t>(std::max<long long>(n, 1)))); }
    ~DevBuf() { cudaFree(p); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

}  // namespace

void coarsen_geom(const dcma_b200_geom &fine, dcma_b200_geom *coarse, int halved[3] /* z, y, x */) {
    *coarse = fine;
    halved[0] = fine.slices >= 8;
    halved[1] = fine.rows >= 8;
    halved[2] = fine.cols >= 8;
    if (halved[0]) {
        coarse->slices = (fine.slices + 1) / 2;
        coarse->pxl_dz = 2.0 * fine.pxl_dz;
    }
    if (halved[1]) {
        coarse->rows = (fine.rows + 1) / 2;
        coarse->pxl_dy = 2.0 * fine.pxl_dy;
    }
    if (halved[2]) {
        coarse->cols = (fine.cols + 1) / 2;
        coarse->pxl_dx = 2.0 * fine.pxl_dx;
    }
}

// low-pass of the level image along the axes that are about to be halved; tmp: scratch of the same size; the result is
// in `img` again
void prefilter_image_device(const dcma_b200_geom &g, float *d_img, float *d_tmp, const int halved[3], cudaStream_t st) {
    const Dim3i n{static_cast<int>(g.slices), static_cast<int>(g.rows), static_cast<int>(g.cols)};
    const long long total = g.slices * g.rows * g.cols * g.channels;
    const int order[3] = {2, 1, 0};  // halved[] is (z, y, x); passes run x, then y, then z
    float *a = d_img, *b = d_tmp;
    for (int pass = 0; pass < 3; ++pass) {
        if (!halved[order[pass]]) continue;
        k_blur_axis<<<grid_for(total), 256, 0, st>>>(a, b, n, static_cast<int>(g.channels), pass);
        DCMA_CUDA(cudaGetLastError());
        std::swap(a, b);
    }
    if (a != d_img) DCMA_CUDA(cudaMemcpyAsync(d_img, a, sizeof(float) * static_cast<size_t>(total), cudaMemcpyDeviceToDevice, st));
}

void restrict_image_device(const dcma_b200_geom &fine, const float *d_fine, const dcma_b200_geom &coarse, const int halved[3],
                           float *d_coarse, cudaStream_t st) {
    const Dim3i nf{static_cast<int>(fine.slices), static_cast<int>(fine.rows), static_cast<int>(fine.cols)};
    const Dim3i nc{static_cast<int>(coarse.slices), static_cast<int>(coarse.rows), static_cast<int>(coarse.cols)};
    const long long total = coarse.slices * coarse.rows * coarse.cols * coarse.channels;
    k_restrict<<<grid_for(total), 256, 0, st>>>(d_fine, d_coarse, nf, nc, static_cast<int>(fine.channels), halved[0], halved[1],
                                                halved[2]);
    DCMA_CUDA(cudaGetLastError());
}

void prolong_field_device(const dcma_b200_geom &coarse, const double *d_coarse_aos, const dcma_b200_geom &fine,
                          const int halved[3], double *d_fine_aos, cudaStream_t st) {
    const Dim3i nf{static_cast<int>(fine.slices), static_cast<int>(fine.rows), static_cast<int>(fine.cols)};
    const Dim3i nc{static_cast<int>(coarse.slices), static_cast<int>(coarse.rows), static_cast<int>(coarse.cols)};
    k_prolong<<<grid_for(fine.slices * fine.rows * fine.cols), 256, 0, st>>>(d_coarse_aos, d_fine_aos, nc, nf, halved[0],
                                                                           halved[1], halved[2]);
    DCMA_CUDA(cudaGetLastError());
}

// host-buffer entry points (tests; also usable on their own)
void restrict_image_host(const dcma_b200_geom &fine, const float *image, dcma_b200_geom *coarse, float *out, int device) {
    if (device >= 0) DCMA_CUDA(cudaSetDevice(device));
    int halved[3];
    coarsen_geom(fine, coarse, halved);
    if (!out) return;  // geometry query
    const long long nf = fine.slices * fine.rows * fine.cols * fine.channels;
    const long long nc = coarse->slices * coarse->rows * coarse->cols * coarse->channels;
    DevBuf<float> df(nf), dc(nc);
    DCMA_CUDA(cudaMemcpy(df.p, image, sizeof(float) * nf, cudaMemcpyHostToDevice));
    restrict_image_device(fine, df.p, *coarse, halved, dc.p, nullptr);
    DCMA_CUDA(cudaMemcpy(out, dc.p, sizeof(float) * nc, cudaMemcpyDeviceToHost));
}

void prolong_field_host(const dcma_b200_geom &fine, const double *coarse_field, double *out, int device) {
    if (device >= 0) DCMA_CUDA(cudaSetDevice(device));
    dcma_b200_geom coarse;
    int halved[3];
    coarsen_geom(fine, &coarse, halved);
    const long long nf = fine.slices * fine.rows * fine.cols, nc = coarse.slices * coarse.rows * coarse.cols;
    DevBuf<double> dc(3 * nc), df(3 * nf);
    DCMA_CUDA(cudaMemcpy(dc.p, coarse_field, sizeof(double) * 3 * nc, cudaMemcpyHostToDevice));
    prolong_field_device(coarse, dc.p, fine, halved, df.p, nullptr);
    DCMA_CUDA(cudaMemcpy(out, df.p, sizeof(double) * 3 * nf, cudaMemcpyDeviceToHost));
}

// The one-shot registration over `levels` levels: H2D of the finest images, restriction on the device, coarse -> fine
// loops with prolongation in between, D2H of the finest field.
void pyramid_register(const dcma_b200_geom &geom, const float *fixed, const float *moving, const dcma_b200_demons_params &params,
                      double *deformation_out, double *mse_history, dcma_b200_run_stats *stats) {
    const int L = params.pyramid_levels;
    if (L < 1 || L > 4) throw std::invalid_argument("pyramid_levels must be in 1..4");
    // n_devices > 1: the FINEST level -- all but 1/7 of the work -- is sharded into z-slabs over devices 0..n-1 of this
    // process (SlabGroup); the coarse levels (1/8, 1/64, ... of the voxels) run whole on the current device
    const bool sharded = params.n_devices > 1;
    if (sharded && L < 2) throw std::invalid_argument("pyramid_register on slabs needs at least two levels");
    if (params.device >= 0) DCMA_CUDA(cudaSetDevice(params.device));
    int dev0 = 0;  // the coarse levels, the restriction and the prolongation run here; slab calls switch devices
    DCMA_CUDA(cudaGetDevice(&dev0));
    std::vector<dcma_b200_geom> g(L);
    std::vector<std::array<int, 3>> hv(L);
    g[0] = geom;
    for (int l = 1; l < L; ++l) coarsen_geom(g[l - 1], &g[l], hv[l].data());
    dcma_b200_demons_params p = params;
    p.pyramid_levels = 1;
    std::vector<std::unique_ptr<EngineBase>> e(L);
    std::unique_ptr<SlabGroup> grp0;
    std::vector<int> devices;
    int64_t halo0 = 0;
    std::unique_ptr<DevBuf<float>> f0, m0;  // sharded: the whole finest images on the current device, for the restriction only
    if (sharded) {
        devices = slab_devices(params.n_devices);
        // the warm start makes |D_z| at the finest level as large as the warp itself: composition gathers reach that far
        halo0 = params.halo_planes > 0 ? params.halo_planes : std::max<int64_t>(SlabGroup::default_halo(g[0], p), 8);
        p.n_devices = 0;
        grp0.reset(new SlabGroup(g[0], p, devices, halo0));
        grp0->load(fixed, moving, false);
        DCMA_CUDA(cudaSetDevice(dev0));
        const long long n0 = g[0].slices * g[0].rows * g[0].cols * g[0].channels;
        f0.reset(new DevBuf<float>(n0));
        m0.reset(new DevBuf<float>(n0));
        DCMA_CUDA(cudaMemcpy(f0->p, fixed, sizeof(float) * static_cast<size_t>(n0), cudaMemcpyHostToDevice));
        DCMA_CUDA(cudaMemcpy(m0->p, moving, sizeof(float) * static_cast<size_t>(n0), cudaMemcpyHostToDevice));
    } else {
        e[0].reset(make_engine(g[0], p, nullptr, SlabSpec{}));
        e[0]->load(fixed, moving, false);
    }
    for (int l = 1; l < L; ++l) {
        const long long nc = g[l].slices * g[l].rows * g[l].cols * g[l].channels;
        DevBuf<float> dF(nc), dM(nc);
        if (e[l - 1]) e[l - 1]->sync();
        {
            // anti-aliasing low-pass of the level l-1 images (copies; the engine keeps the unfiltered ones), then decimation
            const long long nfine = g[l - 1].slices * g[l - 1].rows * g[l - 1].cols * g[l - 1].channels;
            DevBuf<float> lp(nfine), tmp(nfine);
            for (int which = 0; which < 2; ++which) {
                const float *src = (l == 1 && sharded) ? (which == 0 ? f0->p : m0->p)
                                                       : (which == 0 ? e[l - 1]->fixed_dev() : e[l - 1]->moving_dev());
                DCMA_CUDA(cudaMemcpyAsync(lp.p, src, sizeof(float) * static_cast<size_t>(nfine), cudaMemcpyDeviceToDevice, nullptr));
                prefilter_image_device(g[l - 1], lp.p, tmp.p, hv[l].data(), nullptr);
                restrict_image_device(g[l - 1], lp.p, g[l], hv[l].data(), which == 0 ? dF.p : dM.p, nullptr);
            }
        }
        DCMA_CUDA(cudaDeviceSynchronize());
        e[l].reset(make_engine(g[l], p, nullptr, SlabSpec{}));
        e[l]->load(dF.p, dM.p, true);
        e[l]->sync();
        if (l == 1) {
            f0.reset();
            m0.reset();
        }
    }
    dcma_b200_run_stats total;
    std::memset(&total, 0, sizeof(total));
    int64_t hist_off = 0;
    for (int l = L - 1; l >= 0; --l) {
        if (l < L - 1) {
            const long long nc = g[l + 1].slices * g[l + 1].rows * g[l + 1].cols, nf = g[l].slices * g[l].rows * g[l].cols;
            DevBuf<double> aC(3 * nc), aF(3 * nf);
            e[l + 1]->export_deformation_device(aC.p);
            e[l + 1]->sync();
            prolong_field_device(g[l + 1], aC.p, g[l], hv[l + 1].data(), aF.p, nullptr);
            DCMA_CUDA(cudaDeviceSynchronize());
            e[l + 1].reset();
            if (l == 0 && sharded) {
                // finest level on slabs.  A composition gather that leaves the halo fails the run loudly; with the default
                // halo the level is then repeated from the same warm start with a doubled one (as the one-shot call does)
                for (;;) {
                    DCMA_CUDA(cudaSetDevice(dev0));
                    grp0->import_deformation_device(aF.p);
                    dcma_b200_run_stats st;
                    try {
                        grp0->run(params.max_iterations, mse_history ? mse_history + hist_off : nullptr, &st);
                    } catch (const std::logic_error &ex) {
                        const bool overflow = std::string(ex.what()).find("left the slab halo") != std::string::npos;
                        if (!overflow || params.halo_planes > 0 || halo0 * 2 * params.n_devices > geom.slices) throw;
                        halo0 *= 2;
                        grp0.reset();
                        grp0.reset(new SlabGroup(g[0], p, devices, halo0));
                        grp0->load(fixed, moving, false);
                        continue;
                    }
                    hist_off += st.iterations_done;
                    total.iterations_done += st.iterations_done;
                    total.kernel_launches += st.kernel_launches;
                    total.loop_ms += st.loop_ms;
                    break;
                }
                DCMA_CUDA(cudaSetDevice(dev0));
                continue;
            }
            e[l]->import_deformation_device(aF.p);
            e[l]->sync();
        }
        int64_t iters = params.max_iterations;
        if (l >= 1 && l - 1 < 3 && params.pyramid_iterations[l - 1] > 0) iters = params.pyramid_iterations[l - 1];
        dcma_b200_run_stats st;
        e[l]->run(iters, false, mse_history ? mse_history + hist_off : nullptr, &st);
        hist_off += st.iterations_done;
        total.iterations_done += st.iterations_done;
        total.kernel_launches += st.kernel_launches;
        total.loop_ms += st.loop_ms;
    }
    if (sharded)
        grp0->get_deformation(deformation_out);
    else
        e[0]->get(DCMA_B200_BUF_DEFORMATION, deformation_out);
    if (stats) *stats = total;
}

}  // namespace dcma
