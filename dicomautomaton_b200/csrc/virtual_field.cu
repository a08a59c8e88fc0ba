// dcma_b200 -- the two synthetic deformation fields the reference offers for testing registration:
//   GenerateVirtualDataSinusoidalFieldV1   src/Operations/GenerateVirtualDataSinusoidalFieldV1.cc:77-193
//   GenerateVirtualDataVortexFieldV1       src/Operations/GenerateVirtualDataVortexFieldV1.cc:75-178
// Both are two passes over every voxel of a regular stack: (1) an analytic displacement vector per voxel and the extreme
// magnitudes over the volume, (2) a rescale by those extremes.  Here: one kernel per pass, the extremes reduced on the
// device (warp shuffle -> one atomic per block on the order-preserving bit pattern of a non-negative double), the result
// written in the reference's AoS layout.  fp64 throughout; sin/cos/exp are CUDA's (<= 2 ulp from libm), so the result is
// equal to the host loop to ~1e-15 absolute, not bit for bit -- "parity unpinned", checked against the numpy restatement
// oracle/extensions.py::virtual_field_v1.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <stdexcept>

#include "../../include/dcma_b200/demons_c_api.h"
#include "common.cuh"
#include "engine.h"

namespace dcma {
namespace {

struct VF {
    int S, R, C;
    double o[3], ex[3], ey[3], ez[3], dx, dy, dz;
    double b[6];  // x_min, x_max, y_min, y_max, z_min, z_max
    int kind;
};

__device__ __forceinline__ void displacement(const VF &g, const double p[3], double d[3]) {
    const double pi = 3.14159265358979323846;  // == std::acos(-1.0)
    if (g.kind == DCMA_B200_VFIELD_SINUSOIDAL_V1) {
        const double wlx = (g.b[1] - g.b[0]) / 2.0, wly = (g.b[3] - g.b[2]) / 3.0, wlz = (g.b[5] - g.b[4]) / 1.5;  // :77-79
        const double xr = p[0] - g.b[0], yr = p[1] - g.b[2], zr = p[2] - g.b[4];
        const double sx = sin(2.0 * pi * xr / wlx), sy = sin(2.0 * pi * yr / wly), cz = cos(2.0 * pi * zr / wlz);
        d[0] = sy * cz;  // :112-113
        d[1] = sx * cz;  // :116-117
        d[2] = sx * sy;  // :120-121
    } else {
        const double cx = 0.5 * (g.b[0] + g.b[1]), cy = 0.5 * (g.b[2] + g.b[3]), cz = 0.5 * (g.b[4] + g.b[5]);  // (256, 256, 50)
        const double fx = p[0] - cx, fy = p[1] - cy, fz = p[2] - cz;
        const double r = sqrt(fx * fx + fy * fy);
        d[0] = d[1] = d[2] = 0.0;
        if (r > 1e-6) {  // :106
            const double tx = -fy / r, ty = fx / r;
            const double r_max = fmin(cx, cy);
            const double rn = r / r_max;
            const double prof = rn * exp(-2.0 * (rn - 0.5) * (rn - 0.5));  // :120
            const double zn = fabs(fz) / ((g.b[5] - g.b[4]) / 2.0);
            const double zs = 0.3 * exp(-zn * zn);
            d[0] = tx * prof;
            d[1] = ty * prof;
            d[2] = zs * prof * (fz > 0.0 ? 1.0 : -1.0);
        }
    }
}

// order-preserving for non-negative doubles: max/min of the bit patterns == max/min of the values
__global__ void __launch_bounds__(256) k_vfield_fill(const VF g, double *__restrict__ out, unsigned long long *__restrict__ ext) {
    const long long nvox = static_cast<long long>(g.S) * g.R * g.C;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    double mx = 0.0, mn = 1.7976931348623157e308;
    for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvox; v += stride) {
        const int x = static_cast<int>(v % g.C);
        const long long q = v / g.C;
        const int y = static_cast<int>(q % g.R);
        const int z = static_cast<int>(q / g.R);
        double p[3], d[3];
        for (int k = 0; k < 3; ++k) p[k] = g.o[k] + x * g.dx * g.ex[k] + y * g.dy * g.ey[k] + z * g.dz * g.ez[k];
        displacement(g, p, d);
        out[3 * v + 0] = d[0];
        out[3 * v + 1] = d[1];
        out[3 * v + 2] = d[2];
        const double m = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        mx = fmax(mx, m);
        mn = fmin(mn, m);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(ext + 0, static_cast<unsigned long long>(__double_as_longlong(mx)));
        atomicMin(ext + 1, static_cast<unsigned long long>(__double_as_longlong(mn)));
    }
}

__global__ void __launch_bounds__(256) k_vfield_scale(const VF g, double *__restrict__ out, const unsigned long long *__restrict__ ext) {
    const double mx = __longlong_as_double(static_cast<long long>(ext[0])), mn = __longlong_as_double(static_cast<long long>(ext[1]));
    const long long nvox = static_cast<long long>(g.S) * g.R * g.C;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    if (g.kind == DCMA_B200_VFIELD_SINUSOIDAL_V1) {
        if (!(mx > mn && mx > 1e-10)) return;  // :172
        const double range = mx - mn;
        for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvox; v += stride) {
            const double a = out[3 * v], b = out[3 * v + 1], c = out[3 * v + 2];
            const double m = sqrt(a * a + b * b + c * c);
            const double s = (m > 1e-10) ? ((m - mn) / range) / m : 0.0;  // :185
            out[3 * v] = a * s;
            out[3 * v + 1] = b * s;
            out[3 * v + 2] = c * s;
        }
    } else {
        if (!(mx > 1e-10)) return;  // :167
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < 3 * nvox; i += stride) out[i] /= mx;
    }
}

}  // namespace

void virtual_field_host(int kind, const dcma_b200_grid &grid, const double bounds[6], double *out_aos, double extremes[2], int device) {
    if (kind != DCMA_B200_VFIELD_SINUSOIDAL_V1 && kind != DCMA_B200_VFIELD_VORTEX_V1) throw std::invalid_argument("unknown virtual field kind");
    if (grid.geom.slices < 1 || grid.geom.rows < 1 || grid.geom.cols < 1) throw std::invalid_argument("grid extents must be >= 1");
    if (!(bounds[1] > bounds[0]) || !(bounds[3] > bounds[2]) || !(bounds[5] > bounds[4])) throw std::invalid_argument("empty bounds");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw CudaError(cudaErrorNoDevice, "no CUDA device is available (dcma_b200 has no CPU fallback)");
    if (device >= 0) DCMA_CUDA(cudaSetDevice(device));
    VF g;
    g.S = static_cast<int>(grid.geom.slices);
    g.R = static_cast<int>(grid.geom.rows);
    g.C = static_cast<int>(grid.geom.cols);
    g.dx = grid.geom.pxl_dx;
    g.dy = grid.geom.pxl_dy;
    g.dz = grid.geom.pxl_dz;
    for (int i = 0; i < 3; ++i) {
        g.o[i] = grid.origin[i];
        g.ex[i] = grid.ex[i];
        g.ey[i] = grid.ey[i];
        g.ez[i] = grid.ez[i];
    }
    for (int i = 0; i < 6; ++i) g.b[i] = bounds[i];
    g.kind = kind;
    const long long nvox = static_cast<long long>(g.S) * g.R * g.C;
    double *d_out = nullptr;
    unsigned long long *d_ext = nullptr;
    try {
        DCMA_CUDA(cudaMalloc(&d_out, sizeof(double) * 3 * static_cast<size_t>(nvox)));
        DCMA_CUDA(cudaMalloc(&d_ext, 2 * sizeof(unsigned long long)));
        const double init[2] = {0.0, 1.7976931348623157e308};
        DCMA_CUDA(cudaMemcpy(d_ext, init, sizeof init, cudaMemcpyHostToDevice));
        const int blocks = static_cast<int>(std::min<long long>((nvox + 255) / 256, 148LL * 16));
        k_vfield_fill<<<blocks, 256>>>(g, d_out, d_ext);
        DCMA_CUDA(cudaGetLastError());
        k_vfield_scale<<<blocks, 256>>>(g, d_out, d_ext);
        DCMA_CUDA(cudaGetLastError());
        DCMA_CUDA(cudaMemcpy(out_aos, d_out, sizeof(double) * 3 * static_cast<size_t>(nvox), cudaMemcpyDeviceToHost));
        if (extremes) {
            double e[2];
            DCMA_CUDA(cudaMemcpy(e, d_ext, sizeof e, cudaMemcpyDeviceToHost));
            extremes[0] = e[1];  // minimum magnitude before normalisation
            extremes[1] = e[0];  // maximum
        }
    } catch (...) {
        cudaFree(d_out);
        cudaFree(d_ext);
        throw;
    }
    cudaFree(d_out);
    cudaFree(d_ext);
}

}  // namespace dcma
