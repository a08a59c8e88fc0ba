// dcma_b200 -- hand-written sm_100a kernels of the Demons iteration (one per reference SYCL kernel, plus fusions).
//
// Reference kernels (paths relative to /root/reference/):
//   compute_gradient            src/Alignment_Demons.cc:208-272   -> grad_fixed() (recomputed on the fly), k_gradient
//   compute_update_and_mse      src/Alignment_Demons.cc:274-343   -> k_warp_force (+ k_reduce_final)
//   add_update                  src/Alignment_Demons.cc:345-369   -> k_add_update
//   add_update_diffeomorphic    src/Alignment_Demons.cc:371-465   -> k_compose
//   warp_moving                 src/Alignment_Demons.cc:467-550   -> k_warp, and fused into k_warp_force
//   smooth_on_device            src/Alignment_Demons.cc:563-702   -> smooth kernels in smooth_kernels.cuh
//
// HBM layout: images float, interleaved channels exactly as the reference (idx = ((z*R+y)*C+x)*CH+c); vector fields
// are held as THREE COMPONENT VOLUMES (SoA: comp c at base + c*N) of T = float|double, so every field access of a warp
// is one contiguous 128/256-byte row segment.  The reference's AoS double layout exists only at the C-ABI boundary
// (k_export_aos / k_import_aos).
//
// Template parameters: T = field storage type; STRICT = reproduce the reference's x86-64 arithmetic bit for bit
// (true IEEE divisions; the translation unit is compiled with -fmad=false so nothing is contracted).
#pragma once
#include <math_constants.h>

#include "common.cuh"

#ifndef DCMA_NS
#error "define DCMA_NS (one namespace per precision-mode translation unit) before including"
#endif
namespace DCMA_NS {
using namespace ::dcma;

// ------------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------------
template <bool STRICT>
__device__ __forceinline__ double mm_to_index(double d, double p, double ip) {
    if (STRICT) return (p == 1.0) ? d : d / p;  // reference: dx / pxl_dx (cc:403, cc:501)
    return d * ip;
}

struct Tri {
    int x0, y0, z0;
    double tx, ty, tz;
};

// Sample coordinates of voxel (x,y,z) displaced by (dx,dy,dz) mm: cc:401-421 and cc:501-513.
template <bool STRICT>
__device__ __forceinline__ Tri tri_coords(const Geo &g, int x, int y, int z, double dx, double dy, double dz) {
    const double sx = static_cast<double>(x) + mm_to_index<STRICT>(dx, g.px, g.ipx);
    const double sy = static_cast<double>(y) + mm_to_index<STRICT>(dy, g.py, g.ipy);
    const double sz = g.is2d ? static_cast<double>(z) : (static_cast<double>(z) + mm_to_index<STRICT>(dz, g.pz, g.ipz));
    // floor() then clamp into [-2, extent]: both corners of a clamped coordinate are outside the grid either way, and
    // the int conversion stays defined for absurd displacements (the reference's int64 cast would be UB there).
    double fx = floor(sx), fy = floor(sy), fz = floor(sz);
    fx = fmin(fmax(fx, -2.0), static_cast<double>(g.C));
    fy = fmin(fmax(fy, -2.0), static_cast<double>(g.R));
    fz = fmin(fmax(fz, -2.0), static_cast<double>(g.S));
    Tri t;
    t.x0 = static_cast<int>(fx);
    t.y0 = static_cast<int>(fy);
    t.z0 = static_cast<int>(fz);
    t.tx = sx - fx;
    t.ty = sy - fy;
    t.tz = g.is2d ? 0.0 : (sz - fz);
    return t;
}

__device__ __forceinline__ bool finite_f(float v) { return isfinite(v); }

// warp_moving for one channel of one voxel, cc:515-546.  NaN if any corner is outside the grid or non-finite.
__device__ __forceinline__ float warp_one(const float *__restrict__ M, const Geo &g, const Tri &t, int c) {
    const float oob = CUDART_NAN_F;
    const int x1 = t.x0 + 1, y1 = t.y0 + 1, z1 = g.is2d ? t.z0 : (t.z0 + 1);
    if (t.x0 < 0 || x1 >= g.C || t.y0 < 0 || y1 >= g.R || t.z0 < 0 || z1 >= g.S) return oob;
    const long long CH = g.CH;
    const long long row0 = (static_cast<long long>(t.z0) * g.R + t.y0) * g.C;
    const long long row1 = row0 + g.C;  // y1 = y0+1
    const float c000 = __ldg(M + (row0 + t.x0) * CH + c), c100 = __ldg(M + (row0 + x1) * CH + c);
    const float c010 = __ldg(M + (row1 + t.x0) * CH + c), c110 = __ldg(M + (row1 + x1) * CH + c);
    float c001 = c000, c101 = c100, c011 = c010, c111 = c110;  // 2-D: reuse the z0 plane (cc:530-533)
    if (!g.is2d) {
        const long long dz = static_cast<long long>(g.R) * g.C;
        c001 = __ldg(M + (row0 + dz + t.x0) * CH + c);
        c101 = __ldg(M + (row0 + dz + x1) * CH + c);
        c011 = __ldg(M + (row1 + dz + t.x0) * CH + c);
        c111 = __ldg(M + (row1 + dz + x1) * CH + c);
    }
    if (!(finite_f(c000) && finite_f(c100) && finite_f(c010) && finite_f(c110) && finite_f(c001) && finite_f(c101) &&
          finite_f(c011) && finite_f(c111)))
        return oob;
    const double tx = t.tx, ty = t.ty, tz = t.tz;
    const double c00 = static_cast<double>(c000) * (1.0 - tx) + static_cast<double>(c100) * tx;
    const double c10 = static_cast<double>(c010) * (1.0 - tx) + static_cast<double>(c110) * tx;
    const double c01 = static_cast<double>(c001) * (1.0 - tx) + static_cast<double>(c101) * tx;
    const double c11 = static_cast<double>(c011) * (1.0 - tx) + static_cast<double>(c111) * tx;
    const double c0 = c00 * (1.0 - ty) + c10 * ty;
    const double c1 = c01 * (1.0 - ty) + c11 * ty;
    return static_cast<float>(c0 * (1.0 - tz) + c1 * tz);
}

// compute_gradient for one voxel (channel 0 of the fixed image), cc:232-265: central difference in index units,
// one-sided (/1) at the borders, zero if a neighbour is non-finite or the extent is 1.
__device__ __forceinline__ void grad_fixed(const float *__restrict__ F, const Geo &g, int x, int y, int z, double &gx,
                                           double &gy, double &gz) {
    const long long CH = g.CH;
    const long long v = (static_cast<long long>(z) * g.R + y) * g.C + x;
    gx = gy = gz = 0.0;
    if (g.C > 1) {
        const int xl = (x > 0) ? -1 : 0, xr = (x + 1 < g.C) ? 1 : 0;
        const float vl = __ldg(F + (v + xl) * CH), vr = __ldg(F + (v + xr) * CH);
        if (finite_f(vl) && finite_f(vr) && xr != xl)
            gx = (static_cast<double>(vr) - static_cast<double>(vl)) * ((xr - xl == 2) ? 0.5 : 1.0);  // == /(xr-xl), exact
    }
    if (g.R > 1) {
        const int yu = (y > 0) ? -1 : 0, yd = (y + 1 < g.R) ? 1 : 0;
        const float vu = __ldg(F + (v + static_cast<long long>(yu) * g.C) * CH);
        const float vd = __ldg(F + (v + static_cast<long long>(yd) * g.C) * CH);
        if (finite_f(vu) && finite_f(vd) && yd != yu)
            gy = (static_cast<double>(vd) - static_cast<double>(vu)) * ((yd - yu == 2) ? 0.5 : 1.0);
    }
    if (g.S > 1) {
        const long long plane = static_cast<long long>(g.R) * g.C;
        const int zp = (z > 0) ? -1 : 0, zn = (z + 1 < g.S) ? 1 : 0;
        const float vp = __ldg(F + (v + zp * plane) * CH), vn = __ldg(F + (v + zn * plane) * CH);
        if (finite_f(vp) && finite_f(vn) && zn != zp)
            gz = (static_cast<double>(vn) - static_cast<double>(vp)) * ((zn - zp == 2) ? 0.5 : 1.0);
    }
}

// Demons force of one voxel, cc:297-330.  norm_eps = normalization_factor + 1e-10 (same double add as cc:315).
template <bool STRICT>
__device__ __forceinline__ void demons_force(float fixed_val, float moving_val, double gx, double gy, double gz,
                                             double norm_eps, double inv_norm_eps, double max_update, double &ux,
                                             double &uy, double &uz, double &term, int &valid) {
    constexpr double epsilon = 1.0e-10;
    ux = uy = uz = 0.0;
    if (!(finite_f(fixed_val) && finite_f(moving_val))) {
        term = 0.0;
        valid = 0;
        return;
    }
    const double diff = static_cast<double>(fixed_val) - static_cast<double>(moving_val);
    term = diff * diff;
    valid = 1;
    const double gmag_sq = gx * gx + gy * gy + gz * gz;
    const double denom = gmag_sq + (STRICT ? (diff * diff) / norm_eps : (diff * diff) * inv_norm_eps);
    if (denom > epsilon) {
        if (STRICT) {
            ux = diff * gx / denom;
            uy = diff * gy / denom;
            uz = diff * gz / denom;
        } else {
            const double s = diff / denom;
            ux = s * gx;
            uy = s * gy;
            uz = s * gz;
        }
        const double mag = sqrt(ux * ux + uy * uy + uz * uz);
        if (mag > max_update) {
            const double scale = max_update / mag;
            ux *= scale;
            uy *= scale;
            uz *= scale;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_warp_force: warp_moving (of the PREVIOUS iteration, cc:106) fused with compute_update_and_mse (cc:84).
//   w_src = 0: W := M (first iteration: the engine copies M into warped, cc:183 -- NOT a zero-field warp)
//   w_src = 1: W := warp(M, D) computed in registers, never written to HBM
//   w_src = 2: W read from `Wbuf` (stage API: the unfused compute_update_and_mse)
// Tile = 32 x 8 voxels of one slice per 256-thread block, 3-D grid striding over slices; the fixed-image gradient is
// recomputed from F (7-point stencil served by L1/L2) instead of being read back as 3 field components.
// MSE terms are reduced warp-shuffle -> block -> one (sum, count) partial per block (deterministic for a fixed grid).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kTileX = 32, kTileY = 8;

template <typename T, bool STRICT>
__global__ void __launch_bounds__(kTileX *kTileY)
    k_warp_force(const float *__restrict__ F, const float *__restrict__ M, const float *__restrict__ Wbuf,
                 const T *__restrict__ D, T *__restrict__ U, const Geo g, const int w_src, const double norm_eps,
                 const double inv_norm_eps, const double max_update, double *__restrict__ part_sum, long long *__restrict__ part_cnt,
                 const Ctrl *__restrict__ ctrl, const long long it) {
    if (it > ctrl->converged_at) return;
    double acc = 0.0;
    long long cnt = 0;
    const long long N = g.N;
    const int x = blockIdx.x * kTileX + threadIdx.x;
    const int y = blockIdx.y * kTileY + threadIdx.y;
    const bool in_tile = (x < g.C) && (y < g.R);
    for (int z = blockIdx.z; z < g.S; z += gridDim.z) {  // no div/mod: 3-D grid, stride over slices
        if (!in_tile) break;
        const long long v = (static_cast<long long>(z) * g.R + y) * g.C + x;
        const float fixed_val = __ldg(F + v * g.CH);
        float moving_val;
        if (w_src == 0) {
            moving_val = __ldg(M + v * g.CH);
        } else if (w_src == 2) {
            moving_val = __ldg(Wbuf + v * g.CH);
        } else {
            const double dx = static_cast<double>(D[v]), dy = static_cast<double>(D[v + N]),
                         dz = static_cast<double>(D[v + 2 * N]);
            if (!(isfinite(dx) && isfinite(dy) && isfinite(dz))) {  // cc:494-500
                moving_val = CUDART_NAN_F;
            } else {
                const Tri t = tri_coords<STRICT>(g, x, y, z, dx, dy, dz);
                moving_val = warp_one(M, g, t, 0);  // only channel 0 drives the force (cc:294-295)
            }
        }
        double gx, gy, gz;
        grad_fixed(F, g, x, y, z, gx, gy, gz);
        double ux, uy, uz, term;
        int valid;
        demons_force<STRICT>(fixed_val, moving_val, gx, gy, gz, norm_eps, inv_norm_eps, max_update, ux, uy, uz, term,
                             valid);
        U[v] = static_cast<T>(ux);
        U[v + N] = static_cast<T>(uy);
        U[v + 2 * N] = static_cast<T>(uz);
        acc += term;
        cnt += valid;
    }
    // warp-shuffle + block reduction of (sum of squared differences, valid count): cc:334-339 done on the device
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_down_sync(0xffffffffu, acc, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    __shared__ double s_acc[kTileY];
    __shared__ long long s_cnt[kTileY];
    if (threadIdx.x == 0) {
        s_acc[threadIdx.y] = acc;
        s_cnt[threadIdx.y] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        double a = 0.0;
        long long c = 0;
        for (int i = 0; i < kTileY; ++i) {
            a += s_acc[i];
            c += s_cnt[i];
        }
        const int b = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        part_sum[b] = a;
        part_cnt[b] = c;
    }
}

// Final reduction + the loop's stop rule (cc:340-342 and cc:987-994), one block, fixed order => deterministic.
//   count_iter != 0: this FORCE stage belongs to loop iteration `it` (iters_done, history and convergence updated)
__global__ void __launch_bounds__(256)
    k_reduce_final(const double *__restrict__ part_sum, const long long *__restrict__ part_cnt, const int nparts,
                   Ctrl *__restrict__ ctrl, double *__restrict__ mse_hist, const long long it, const long long hist_idx,
                   const int count_iter) {
    if (it > ctrl->converged_at) return;
    __shared__ double s_acc[256];
    __shared__ long long s_cnt[256];
    double a = 0.0;
    long long c = 0;
    for (int i = threadIdx.x; i < nparts; i += 256) {
        a += part_sum[i];
        c += part_cnt[i];
    }
    s_acc[threadIdx.x] = a;
    s_cnt[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_acc[threadIdx.x] += s_acc[threadIdx.x + o];
            s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double mse = s_acc[0];
        const long long n = s_cnt[0];
        if (n > 0) mse /= static_cast<double>(n);
        ctrl->last_mse = mse;
        ctrl->last_n = n;
        if (count_iter) {
            if (mse_hist) mse_hist[hist_idx] = mse;
            ctrl->iters_done += 1;
            const double change = fabs(ctrl->prev_mse - mse);
            if (change < ctrl->threshold && it > 0) {
                ctrl->converged_at = it;  // this iteration's update is still applied (the reference breaks after it)
            } else {
                ctrl->prev_mse = mse;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_warp: warp_moving as its own kernel (stage API, KeepDeformedImages, dcma_b200_warp_image): all channels.
// D may be SoA T (engine) -- AoS double input goes through k_import_aos first.
// ------------------------------------------------------------------------------------------------------------------
template <typename T, bool STRICT>
__global__ void __launch_bounds__(kTileX *kTileY)
    k_warp(const float *__restrict__ M, const T *__restrict__ D, float *__restrict__ W, const Geo g, const int use_oob,
           const float oob_value) {
    const long long N = g.N;
    const int x = blockIdx.x * kTileX + threadIdx.x;
    const int y = blockIdx.y * kTileY + threadIdx.y;
    const bool in_tile = (x < g.C) && (y < g.R);
    for (int z = blockIdx.z; z < g.S; z += gridDim.z) {  // no div/mod: 3-D grid, stride over slices
        if (!in_tile) break;
        const long long v = (static_cast<long long>(z) * g.R + y) * g.C + x;
        const double dx = static_cast<double>(D[v]), dy = static_cast<double>(D[v + N]),
                     dz = static_cast<double>(D[v + 2 * N]);
        const bool ok = isfinite(dx) && isfinite(dy) && isfinite(dz);
        Tri t;
        if (ok) t = tri_coords<STRICT>(g, x, y, z, dx, dy, dz);
        for (int c = 0; c < g.CH; ++c) {
            float w = ok ? warp_one(M, g, t, c) : CUDART_NAN_F;
            if (use_oob && !finite_f(w)) w = oob_value;
            W[v * g.CH + c] = w;
        }
    }
}

// k_gradient: compute_gradient as its own kernel, AoS double output (stage API / inspection only).
__global__ void __launch_bounds__(kTileX *kTileY)
    k_gradient(const float *__restrict__ F, double *__restrict__ G_aos, const Geo g) {
    const int x = blockIdx.x * kTileX + threadIdx.x;
    const int y = blockIdx.y * kTileY + threadIdx.y;
    const bool in_tile = (x < g.C) && (y < g.R);
    for (int z = blockIdx.z; z < g.S; z += gridDim.z) {  // no div/mod: 3-D grid, stride over slices
        if (!in_tile) break;
        const long long v = (static_cast<long long>(z) * g.R + y) * g.C + x;
        double gx, gy, gz;
        grad_fixed(F, g, x, y, z, gx, gy, gz);
        G_aos[v * 3 + 0] = gx;
        G_aos[v * 3 + 1] = gy;
        G_aos[v * 3 + 2] = gz;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_add_update: D[c] += U[c] * pxl_c   (cc:361-366)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    k_add_update(T *__restrict__ D, const T *__restrict__ U, const long long N, const double px, const double py,
                 const double pz, const Ctrl *__restrict__ ctrl, const long long it) {
    if (it > ctrl->converged_at) return;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < 3 * N; i += stride) {
        const double p = (i < N) ? px : ((i < 2 * N) ? py : pz);
        D[i] = static_cast<T>(static_cast<double>(D[i]) + static_cast<double>(U[i]) * p);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_compose: add_update_diffeomorphic (cc:371-465):  D(p) += U(p + D(p)/pxl) * pxl, trilinear, ZERO outside the grid,
// no NaN checks; the four upper-z corners are narrowed to float exactly as the reference does (cc:439-442).
// Reads D only at self (in place is race-free), gathers U at 8 corners x 3 components.
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ double fetch_u(const T *__restrict__ Uc, const Geo &g, int zz, int yy, int xx) {
    if (xx < 0 || xx >= g.C || yy < 0 || yy >= g.R || zz < 0 || zz >= g.S) return 0.0;
    return static_cast<double>(Uc[(static_cast<long long>(zz) * g.R + yy) * g.C + xx]);
}

template <typename T, bool STRICT>
__global__ void __launch_bounds__(kTileX *kTileY)
    k_compose(T *__restrict__ D, const T *__restrict__ U, const Geo g, const Ctrl *__restrict__ ctrl, const long long it) {
    if (it > ctrl->converged_at) return;
    const long long N = g.N;
    const int x = blockIdx.x * kTileX + threadIdx.x;
    const int y = blockIdx.y * kTileY + threadIdx.y;
    const bool in_tile = (x < g.C) && (y < g.R);
    for (int z = blockIdx.z; z < g.S; z += gridDim.z) {  // no div/mod: 3-D grid, stride over slices
        if (!in_tile) break;
        const long long v = (static_cast<long long>(z) * g.R + y) * g.C + x;
        const double dx = static_cast<double>(D[v]), dy = static_cast<double>(D[v + N]),
                     dz = static_cast<double>(D[v + 2 * N]);
        const Tri t = tri_coords<STRICT>(g, x, y, z, dx, dy, dz);
        const int x1 = t.x0 + 1, y1 = t.y0 + 1, z1 = g.is2d ? t.z0 : (t.z0 + 1);
        const double tx = t.tx, ty = t.ty, tz = t.tz;
        double u[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const T *Uc = U + c * N;
            const double c000 = fetch_u(Uc, g, t.z0, t.y0, t.x0), c100 = fetch_u(Uc, g, t.z0, t.y0, x1);
            const double c010 = fetch_u(Uc, g, t.z0, y1, t.x0), c110 = fetch_u(Uc, g, t.z0, y1, x1);
            float c001, c101, c011, c111;
            if (g.is2d) {
                c001 = static_cast<float>(c000);
                c101 = static_cast<float>(c100);
                c011 = static_cast<float>(c010);
                c111 = static_cast<float>(c110);
            } else {
                c001 = static_cast<float>(fetch_u(Uc, g, z1, t.y0, t.x0));
                c101 = static_cast<float>(fetch_u(Uc, g, z1, t.y0, x1));
                c011 = static_cast<float>(fetch_u(Uc, g, z1, y1, t.x0));
                c111 = static_cast<float>(fetch_u(Uc, g, z1, y1, x1));
            }
            const double c00 = c000 * (1.0 - tx) + c100 * tx;
            const double c10 = c010 * (1.0 - tx) + c110 * tx;
            const double c01 = static_cast<double>(c001) * (1.0 - tx) + static_cast<double>(c101) * tx;
            const double c11 = static_cast<double>(c011) * (1.0 - tx) + static_cast<double>(c111) * tx;
            const double c0 = c00 * (1.0 - ty) + c10 * ty;
            const double c1 = c01 * (1.0 - ty) + c11 * ty;
            u[c] = c0 * (1.0 - tz) + c1 * tz;
        }
        D[v] = static_cast<T>(dx + u[0] * g.px);
        D[v + N] = static_cast<T>(dy + u[1] * g.py);
        D[v + 2 * N] = static_cast<T>(dz + u[2] * g.pz);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// boundary conversions: reference AoS double [z][y][x][3]  <->  device SoA T [3][z][y][x]
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_export_aos(const T *__restrict__ soa, double *__restrict__ aos, const long long N) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < 3 * N; i += stride) {
        const long long v = i / 3;
        const int c = static_cast<int>(i - v * 3);
        aos[i] = static_cast<double>(soa[v + c * N]);  // coalesced write, 3-way strided (cached) read
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_import_aos(const double *__restrict__ aos, T *__restrict__ soa, const long long N) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < 3 * N; i += stride) {
        const long long v = i / 3;
        const int c = static_cast<int>(i - v * 3);
        soa[v + c * N] = static_cast<T>(aos[i]);
    }
}

}  // namespace DCMA_NS
