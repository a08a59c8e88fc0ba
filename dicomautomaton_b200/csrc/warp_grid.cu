// dcma_b200 -- warp of an image that lives on ANOTHER regular grid than the deformation field:
// the per-voxel body of the reference's WarpImages for a deformation_field transform,
//   ref_p  = img.position(row, col)                                  src/Operations/WarpImages.cc:424
//   corr_p = ref_p + D(ref_p)      (three trilinear field lookups)   src/Alignment_Field.cc:138-153, WarpImages.cc:442-443
//   voxel  = img_adj.trilinearly_interpolate(corr_p, chan, 0.0)      src/Operations/WarpImages.cc:452 (out of bounds -> 0.0, :197)
// e.g. a co-registered RT dose grid (2.5 mm) warped with a field estimated on the CT grid (1 mm), BASELINE config 3.
// The trilinear lookups of the reference are Ygor's planar_image_adjacency (not vendored, unpinned); the definition
// used here, restated on the CPU in oracle/extensions.py: index coordinates u = ((p - origin) . e_axis) / spacing; a
// point is inside a grid iff 0 <= u <= n-1 on every axis (a single-plane axis: |u| <= 1/2, no interpolation along it);
// lower corner i0 = min(floor(u), n-2); any non-finite corner makes the sample non-finite; a non-finite field sample
// or image sample yields `oob_value`.  fp64 throughout, like the reference.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <stdexcept>

#include "../../include/dcma_b200/demons_c_api.h"
#include "common.cuh"
#include "engine.h"

namespace dcma {
namespace {

struct GridK {
    int S, R, C, CH;
    double o[3], ex[3], ey[3], ez[3], dx, dy, dz;
};

__device__ __forceinline__ bool axis_setup(double u, int n, int &i0, int &i1, double &t) {
    if (n == 1) {  // a single plane along this axis: inside its own thickness, no interpolation
        i0 = i1 = 0;
        t = 0.0;
        return fabs(u) <= 0.5;
    }
    if (!(u >= 0.0 && u <= static_cast<double>(n - 1))) return false;  // also rejects NaN
    i0 = min(static_cast<int>(floor(u)), n - 2);
    i1 = i0 + 1;
    t = u - static_cast<double>(i0);
    return true;
}

__device__ __forceinline__ bool to_index(const GridK &g, const double p[3], double u[3]) {
    const double q[3] = {p[0] - g.o[0], p[1] - g.o[1], p[2] - g.o[2]};
    u[0] = (q[0] * g.ex[0] + q[1] * g.ex[1] + q[2] * g.ex[2]) / g.dx;
    u[1] = (q[0] * g.ey[0] + q[1] * g.ey[1] + q[2] * g.ey[2]) / g.dy;
    u[2] = (q[0] * g.ez[0] + q[1] * g.ez[1] + q[2] * g.ez[2]) / g.dz;
    return true;
}

template <typename V>
__device__ __forceinline__ double trilinear(const V *__restrict__ data, const GridK &g, int nch, int c, const double u[3], bool &ok) {
    int x0, x1, y0, y1, z0, z1;
    double tx, ty, tz;
    ok = axis_setup(u[0], g.C, x0, x1, tx) & axis_setup(u[1], g.R, y0, y1, ty) & axis_setup(u[2], g.S, z0, z1, tz);
    if (!ok) return 0.0;
    auto at = [&](int z, int y, int x) {
        return static_cast<double>(data[((static_cast<long long>(z) * g.R + y) * g.C + x) * nch + c]);
    };
    const double c00 = at(z0, y0, x0) * (1.0 - tx) + at(z0, y0, x1) * tx;
    const double c10 = at(z0, y1, x0) * (1.0 - tx) + at(z0, y1, x1) * tx;
    const double c01 = at(z1, y0, x0) * (1.0 - tx) + at(z1, y0, x1) * tx;
    const double c11 = at(z1, y1, x0) * (1.0 - tx) + at(z1, y1, x1) * tx;
    const double c0 = c00 * (1.0 - ty) + c10 * ty;
    const double c1 = c01 * (1.0 - ty) + c11 * ty;
    const double v = c0 * (1.0 - tz) + c1 * tz;
    ok = isfinite(v);
    return v;
}

__global__ void __launch_bounds__(256)
    k_warp_grid(const GridK fg, const double *__restrict__ field_aos, const GridK ig, const float *__restrict__ image,
                const unsigned char *__restrict__ mask, const float oob_value, float *__restrict__ out) {
    const long long nvox = static_cast<long long>(ig.S) * ig.R * ig.C;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvox; v += stride) {
        if (mask && !mask[v]) {  // outside the region of interest: voxel left as it is (WarpImages mutates ROI voxels only)
            for (int c = 0; c < ig.CH; ++c) out[v * ig.CH + c] = image[v * ig.CH + c];
            continue;
        }
        const int x = static_cast<int>(v % ig.C);
        const long long r = v / ig.C;
        const int y = static_cast<int>(r % ig.R);
        const int z = static_cast<int>(r / ig.R);
        double p[3];
        for (int k = 0; k < 3; ++k) p[k] = ig.o[k] + x * ig.dx * ig.ex[k] + y * ig.dy * ig.ey[k] + z * ig.dz * ig.ez[k];
        double u[3];
        to_index(fg, p, u);
        bool okx, oky, okz;
        const double ddx = trilinear(field_aos, fg, 3, 0, u, okx);
        const double ddy = trilinear(field_aos, fg, 3, 1, u, oky);
        const double ddz = trilinear(field_aos, fg, 3, 2, u, okz);
        const bool okd = okx && oky && okz;
        // displacement components are along the field's index axes (x = column direction, ...), in mm
        double cp[3];
        for (int k = 0; k < 3; ++k) cp[k] = p[k] + ddx * fg.ex[k] + ddy * fg.ey[k] + ddz * fg.ez[k];
        double ui[3];
        to_index(ig, cp, ui);
        for (int c = 0; c < ig.CH; ++c) {
            bool ok = false;
            const double val = okd ? trilinear(image, ig, ig.CH, c, ui, ok) : 0.0;
            out[v * ig.CH + c] = (okd && ok) ? static_cast<float>(val) : oob_value;
        }
    }
}

// The same body with THREE grids, as the operation has them (WarpImages.cc:385-456): the voxels of the OUTPUT stack (a copy
// of the ReferenceImageSelection, the geometry placeholder) are visited, displaced through the field, and take the
// trilinear sample of the SOURCE stack (the ImageSelection).  `out` holds the placeholder's voxels on entry: voxels
// outside the mask, and channels other than `channel` (>= 0), keep them -- the operation's functor only assigns
// `voxel_val` for bounded voxels of the selected channel (:419-422, 454).
__global__ void __launch_bounds__(256)
    k_warp_to_grid(const GridK fg, const double *__restrict__ field_aos, const GridK sg, const float *__restrict__ source,
                   const GridK og, const unsigned char *__restrict__ mask, const int channel, const float oob_value,
                   float *__restrict__ out) {
    const long long nvox = static_cast<long long>(og.S) * og.R * og.C;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvox; v += stride) {
        if (mask && !mask[v]) continue;
        const int x = static_cast<int>(v % og.C);
        const long long r = v / og.C;
        const int y = static_cast<int>(r % og.R);
        const int z = static_cast<int>(r / og.R);
        double p[3];
        for (int k = 0; k < 3; ++k) p[k] = og.o[k] + x * og.dx * og.ex[k] + y * og.dy * og.ey[k] + z * og.dz * og.ez[k];
        double u[3];
        to_index(fg, p, u);
        bool okx, oky, okz;
        const double ddx = trilinear(field_aos, fg, 3, 0, u, okx);
        const double ddy = trilinear(field_aos, fg, 3, 1, u, oky);
        const double ddz = trilinear(field_aos, fg, 3, 2, u, okz);
        const bool okd = okx && oky && okz;
        double cp[3];
        for (int k = 0; k < 3; ++k) cp[k] = p[k] + ddx * fg.ex[k] + ddy * fg.ey[k] + ddz * fg.ez[k];
        double us[3];
        to_index(sg, cp, us);
        for (int c = 0; c < og.CH; ++c) {
            if (channel >= 0 && c != channel) continue;
            bool ok = false;
            const double val = okd ? trilinear(source, sg, sg.CH, c, us, ok) : 0.0;
            out[v * og.CH + c] = (okd && ok) ? static_cast<float>(val) : oob_value;
        }
    }
}

// resample_image_to_reference_grid (src/Alignment_Demons.cc:709-755): every voxel of the destination grid takes the
// trilinear sample of the source image at its own position, `oob_value` (the reference: NaN, cc:743) outside.
__global__ void __launch_bounds__(256)
    k_resample_grid(const GridK sg, const float *__restrict__ image, const GridK dg, const float oob_value, float *__restrict__ out) {
    const long long nvox = static_cast<long long>(dg.S) * dg.R * dg.C;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvox; v += stride) {
        const int x = static_cast<int>(v % dg.C);
        const long long r = v / dg.C;
        const int y = static_cast<int>(r % dg.R);
        const int z = static_cast<int>(r / dg.R);
        double p[3], u[3];
        for (int k = 0; k < 3; ++k) p[k] = dg.o[k] + x * dg.dx * dg.ex[k] + y * dg.dy * dg.ey[k] + z * dg.dz * dg.ez[k];
        to_index(sg, p, u);
        for (int c = 0; c < sg.CH; ++c) {
            bool ok = false;
            const double val = trilinear(image, sg, sg.CH, c, u, ok);
            out[v * sg.CH + c] = ok ? static_cast<float>(val) : oob_value;
        }
    }
}

GridK make_gridk(const dcma_b200_grid &g) {
    if (g.geom.slices < 1 || g.geom.rows < 1 || g.geom.cols < 1 || g.geom.channels < 1)
        throw std::invalid_argument("grid extents and channel count must be >= 1");
    if (!(g.geom.pxl_dx > 0.0) || !(g.geom.pxl_dy > 0.0) || !(g.geom.pxl_dz > 0.0)) throw std::invalid_argument("grid spacing must be > 0");
    GridK k;
    k.S = static_cast<int>(g.geom.slices);
    k.R = static_cast<int>(g.geom.rows);
    k.C = static_cast<int>(g.geom.cols);
    k.CH = static_cast<int>(g.geom.channels);
    k.dx = g.geom.pxl_dx;
    k.dy = g.geom.pxl_dy;
    k.dz = g.geom.pxl_dz;
    for (int i = 0; i < 3; ++i) {
        k.o[i] = g.origin[i];
        k.ex[i] = g.ex[i];
        k.ey[i] = g.ey[i];
        k.ez[i] = g.ez[i];
    }
    auto norm = [](const double *a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
    if (std::abs(norm(k.ex) - 1.0) > 1e-6 || std::abs(norm(k.ey) - 1.0) > 1e-6 || std::abs(norm(k.ez) - 1.0) > 1e-6)
        throw std::invalid_argument("grid axes must be unit vectors");
    return k;
}

}  // namespace

void warp_image_grid_host(const dcma_b200_grid &field_grid, const double *deformation_aos, const dcma_b200_grid &image_grid,
                          const float *image, const unsigned char *mask, float oob_value, float *out, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw CudaError(cudaErrorNoDevice, "no CUDA device is available (dcma_b200 has no CPU fallback)");
    if (device >= 0) DCMA_CUDA(cudaSetDevice(device));
    GridK fg = make_gridk(field_grid);
    const GridK ig = make_gridk(image_grid);
    fg.CH = 3;
    const long long nf = static_cast<long long>(fg.S) * fg.R * fg.C, ni = static_cast<long long>(ig.S) * ig.R * ig.C;
    double *dD = nullptr;
    float *dI = nullptr, *dO = nullptr;
    unsigned char *dMask = nullptr;
    auto cleanup = [&]() {
        cudaFree(dD);
        cudaFree(dI);
        cudaFree(dO);
        cudaFree(dMask);
    };
    try {
        DCMA_CUDA(cudaMalloc(&dD, sizeof(double) * 3 * nf));
        DCMA_CUDA(cudaMalloc(&dI, sizeof(float) * ni * ig.CH));
        DCMA_CUDA(cudaMalloc(&dO, sizeof(float) * ni * ig.CH));
        DCMA_CUDA(cudaMemcpy(dD, deformation_aos, sizeof(double) * 3 * nf, cudaMemcpyHostToDevice));
        DCMA_CUDA(cudaMemcpy(dI, image, sizeof(float) * ni * ig.CH, cudaMemcpyHostToDevice));
        if (mask) {
            DCMA_CUDA(cudaMalloc(&dMask, static_cast<size_t>(ni)));
            DCMA_CUDA(cudaMemcpy(dMask, mask, static_cast<size_t>(ni), cudaMemcpyHostToDevice));
        }
        const int grid = static_cast<int>(std::min<long long>((ni + 255) / 256, 148LL * 16));
        k_warp_grid<<<grid, 256>>>(fg, dD, ig, dI, dMask, oob_value, dO);
        DCMA_CUDA(cudaGetLastError());
        DCMA_CUDA(cudaMemcpy(out, dO, sizeof(float) * ni * ig.CH, cudaMemcpyDeviceToHost));
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}

void warp_to_grid_host(const dcma_b200_grid &field_grid, const double *deformation_aos, const dcma_b200_grid &source_grid,
                       const float *source, const dcma_b200_grid &out_grid, const unsigned char *mask, int channel, float oob_value,
                       float *out_inout, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw CudaError(cudaErrorNoDevice, "no CUDA device is available (dcma_b200 has no CPU fallback)");
    if (device >= 0) DCMA_CUDA(cudaSetDevice(device));
    GridK fg = make_gridk(field_grid);
    const GridK sg = make_gridk(source_grid), og = make_gridk(out_grid);
    fg.CH = 3;
    if (channel >= og.CH) throw std::invalid_argument("Encountered image without requested channel");
    if ((channel < 0 ? og.CH : channel + 1) > sg.CH) throw std::invalid_argument("the source stack lacks a channel of the output stack");
    const long long nf = static_cast<long long>(fg.S) * fg.R * fg.C, ns = static_cast<long long>(sg.S) * sg.R * sg.C * sg.CH,
                    no = static_cast<long long>(og.S) * og.R * og.C;
    double *dD = nullptr;
    float *dS = nullptr, *dO = nullptr;
    unsigned char *dMask = nullptr;
    auto cleanup = [&]() {
        cudaFree(dD);
        cudaFree(dS);
        cudaFree(dO);
        cudaFree(dMask);
    };
    try {
        DCMA_CUDA(cudaMalloc(&dD, sizeof(double) * 3 * nf));
        DCMA_CUDA(cudaMalloc(&dS, sizeof(float) * ns));
        DCMA_CUDA(cudaMalloc(&dO, sizeof(float) * no * og.CH));
        DCMA_CUDA(cudaMemcpy(dD, deformation_aos, sizeof(double) * 3 * nf, cudaMemcpyHostToDevice));
        DCMA_CUDA(cudaMemcpy(dS, source, sizeof(float) * ns, cudaMemcpyHostToDevice));
        DCMA_CUDA(cudaMemcpy(dO, out_inout, sizeof(float) * no * og.CH, cudaMemcpyHostToDevice));
        if (mask) {
            DCMA_CUDA(cudaMalloc(&dMask, static_cast<size_t>(no)));
            DCMA_CUDA(cudaMemcpy(dMask, mask, static_cast<size_t>(no), cudaMemcpyHostToDevice));
        }
        const int grid = static_cast<int>(std::min<long long>((no + 255) / 256, 148LL * 16));
        k_warp_to_grid<<<grid, 256>>>(fg, dD, sg, dS, og, dMask, channel, oob_value, dO);
        DCMA_CUDA(cudaGetLastError());
        DCMA_CUDA(cudaMemcpy(out_inout, dO, sizeof(float) * no * og.CH, cudaMemcpyDeviceToHost));
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}

void resample_to_grid_device(const dcma_b200_grid &src_grid, const float *d_image, const dcma_b200_grid &dst_grid, float oob_value,
                             float *d_out, void *stream) {
    const GridK sg = make_gridk(src_grid);
    const GridK dg = make_gridk(dst_grid);
    const long long nv = static_cast<long long>(dg.S) * dg.R * dg.C;
    k_resample_grid<<<static_cast<int>(std::min<long long>((nv + 255) / 256, 148LL * 16)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        sg, d_image, dg, oob_value, d_out);
    DCMA_CUDA(cudaGetLastError());
}

void resample_to_grid_host(const dcma_b200_grid &src_grid, const float *image, const dcma_b200_grid &dst_grid, float oob_value,
                           float *out, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw CudaError(cudaErrorNoDevice, "no CUDA device is available (dcma_b200 has no CPU fallback)");
    if (device >= 0) DCMA_CUDA(cudaSetDevice(device));
    const GridK sg = make_gridk(src_grid);
    const GridK dg = make_gridk(dst_grid);
    const long long ns = static_cast<long long>(sg.S) * sg.R * sg.C * sg.CH, nd = static_cast<long long>(dg.S) * dg.R * dg.C * sg.CH;
    float *dI = nullptr, *dO = nullptr;
    try {
        DCMA_CUDA(cudaMalloc(&dI, sizeof(float) * ns));
        DCMA_CUDA(cudaMalloc(&dO, sizeof(float) * nd));
        DCMA_CUDA(cudaMemcpy(dI, image, sizeof(float) * ns, cudaMemcpyHostToDevice));
        const long long nv = static_cast<long long>(dg.S) * dg.R * dg.C;
        k_resample_grid<<<static_cast<int>(std::min<long long>((nv + 255) / 256, 148LL * 16)), 256>>>(sg, dI, dg, oob_value, dO);
        DCMA_CUDA(cudaGetLastError());
        DCMA_CUDA(cudaMemcpy(out, dO, sizeof(float) * nd, cudaMemcpyDeviceToHost));
    } catch (...) {
        cudaFree(dI);
        cudaFree(dO);
        throw;
    }
    cudaFree(dI);
    cudaFree(dO);
}

}  // namespace dcma
