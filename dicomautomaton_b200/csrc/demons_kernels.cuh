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
#include <type_traits>
#include <math_constants.h>

#include "common.cuh"

#ifndef DCMA_NS
#error "define DCMA_NS (one namespace per precision-mode translation unit) before including"
#endif
namespace DCMA_NS {
using namespace ::dcma;

constexpr int kTileX = 32, kTileY = 8;  // per-voxel kernels: 32 x 8 voxels of one slice per 256-thread block

// The per-voxel kernels march along z and every step ends in a chain  D -> sample position -> gathers -> arithmetic,
// whose DRAM latency (~1000 SM cycles loaded) the resident warps cannot cover (ncu: >70 % long-scoreboard stalls).
// Each step therefore asks for the lines it will need kPrefetchAhead slices later -- the voxel's own row of every
// streamed array, which is where the gathers land for displacements of a few voxels -- so that they come from L2.
#ifndef DCMA_PREFETCH_AHEAD
#define DCMA_PREFETCH_AHEAD 3
#endif
constexpr int kPrefetchAhead = DCMA_PREFETCH_AHEAD;
#ifndef DCMA_PREFETCH_L1
#define DCMA_PREFETCH_L1 0
#endif
__device__ __forceinline__ void prefetch_line(const void *p) {
#if DCMA_PREFETCH_L1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// scalar helpers.  B200 runs F2F/F2I/I2F/FRND/MUFU at 16 lanes/clk/SM (measured with ncu: the "xu" pipe), four times
// slower than the fp64 pipe, so floor() and int<->double conversions are done with fp64/fp32 adds instead.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool finite_f(float v) { return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u; }
__device__ __forceinline__ bool finite_d(double v) { return (static_cast<unsigned>(__double2hiint(v)) & 0x7ff00000u) != 0x7ff00000u; }

// floor of |s| < 2^31 as (double f, int i), exactly: the round-DOWN add  s + 1.5*2^52  is  1.5*2^52 + floor(s)  (ulp 1 there),
// whose two's complement low word is floor(s); subtracting the magic back is exact.  No compare, no conversion pipe.
__device__ __forceinline__ void floor_split(double s, double &f, int &i) {
    constexpr double kMagic = 6755399441055744.0;
    const double t = __dadd_rd(s, kMagic);
    f = __dadd_rn(t, -kMagic);
    i = __double2loint(t);
}
// same for |s| < 2^22 in fp32 (magic 1.5*2^23 = 0x4B400000: consecutive integers are consecutive bit patterns)
__device__ __forceinline__ void floor_split(float s, float &f, int &i) {
    constexpr float kMagic = 12582912.0f;
    const float t = __fadd_rd(s, kMagic);
    f = __fadd_rn(t, -kMagic);
    i = __float_as_int(t) - 0x4B400000;
}

// x rounded to float precision and widened again (the reference narrows four corners to float, cc:439-442).
//   STRICT: the real conversions.  Otherwise Veltkamp's split (3 fp64 ops, no conversion pipe): identical except on
//   exact ties and outside the float exponent range, neither of which matters below 1e-12 mm.
template <bool STRICT>
__device__ __forceinline__ double narrow_to_float(double x) {
    if (STRICT) return static_cast<double>(static_cast<float>(x));
    const double c = __dmul_rn(x, 536870913.0);  // 2^29 + 1
    return __dsub_rn(c, __dsub_rn(c, x));
}
template <bool STRICT>
__device__ __forceinline__ float narrow_to_float(float x) {
    return x;
}

// (double)a - (double)b, or 0 if either value is non-finite (cc:238, 249, 260).
__device__ __forceinline__ double diff_exact(float a, float b) {
    const double d = static_cast<double>(a) - static_cast<double>(b);
    return (finite_f(a) && finite_f(b)) ? d : 0.0;
}

// ------------------------------------------------------------------------------------------------------------------
// Sample position of a voxel displaced by d mm along one axis (cc:401-421, cc:501-513): integer corner i0, fraction t
// and the coordinate magnitude |s| (for the range guard of tri_coords).  Nothing is clamped: a coordinate that is
// NaN, infinite or beyond the exact range of floor_split makes Tri::ok false, and such a sample point has every
// corner outside the grid anyway (extents are < 2^20 per axis; the reference's int64 cast would be undefined there).
// ------------------------------------------------------------------------------------------------------------------
template <bool STRICT>
__device__ __forceinline__ void axis_coord(int xi, double xd, double d, double p, double ip, int &i0, double &t,
                                           double &mag) {
    (void)xi;
    const double s = xd + (STRICT ? ((p == 1.0) ? d : d / p) : d * ip);  // reference: x + dx / pxl_dx
    double f;
    floor_split(s, f, i0);
    t = s - f;
    mag = fabs(s);
}
template <bool STRICT>
__device__ __forceinline__ void axis_coord(int xi, float xd, float d, float p, float ip, int &i0, float &t, float &mag) {
    (void)xd;
    (void)p;
    // fp32 mode: never add the voxel index to the displacement (that would cost 10 bits of the fraction)
    const float s = d * ip;
    float f;
    int n;
    floor_split(s, f, n);
    i0 = xi + n;
    t = s - f;
    mag = fabsf(s);
}

// per-kernel constants of the compute type
template <typename CT>
struct AxisK {
    CT px, py, pz, ipx, ipy, ipz;
    __device__ __forceinline__ explicit AxisK(const Geo &g)
        : px(static_cast<CT>(g.px)), py(static_cast<CT>(g.py)), pz(static_cast<CT>(g.pz)),
          ipx(static_cast<CT>(g.ipx)), ipy(static_cast<CT>(g.ipy)), ipz(static_cast<CT>(g.ipz)) {}
};

template <typename CT>
struct Tri {
    int x0, y0, z0;
    CT tx, ty, tz;
    bool ok;  // coordinates finite and inside floor_split's exact range; false => every corner is outside the grid
};

template <typename CT, bool STRICT>
__device__ __forceinline__ Tri<CT> tri_coords(const Geo &g, const AxisK<CT> &k, int x, int y, int z, CT xd, CT yd, CT zd,
                                              CT dx, CT dy, CT dz) {
    Tri<CT> t;
    CT mx, my, mz;
    axis_coord<STRICT>(x, xd, dx, k.px, k.ipx, t.x0, t.tx, mx);
    axis_coord<STRICT>(y, yd, dy, k.py, k.ipy, t.y0, t.ty, my);
    if (g.is2d) {  // cc:405, 413, 421 / cc:503, 510, 513
        t.z0 = z;
        t.tz = CT(0);
        mz = CT(0);
    } else {
        axis_coord<STRICT>(z, zd, dz, k.pz, k.ipz, t.z0, t.tz, mz);
    }
    // one compare for all three axes; NaN compares false.  2^21 (float: |s| < 2^22 needed) / 2^30 (double: < 2^31)
    t.ok = (mx + my + mz) < (sizeof(CT) == 4 ? CT(2097152.0) : CT(1073741824.0));
    return t;
}

// ------------------------------------------------------------------------------------------------------------------
// warp_moving for one channel of one voxel (cc:515-546): NaN if any corner is outside the grid or non-finite.
// Interpolating first and testing the RESULT is equivalent to the reference's eight isfinite tests: with weights in
// [0,1] finite corners give a finite value inside the float range, and a non-finite corner always yields NaN/Inf.
//   double path: the reference's fp64 arithmetic, rounded to float once (cc:546).
//   float  path: fp32 on differences to the first corner (error ~1e-6 HU instead of ~1e-4 HU for plain fp32).
// Returns diff = fixed - warped through `diff` when WANT_DIFF, and the warped value itself.
// ------------------------------------------------------------------------------------------------------------------
// Grid limits of the "all eight corners inside" test (cc:516-519) as three unsigned compares.
struct InsideK {
    unsigned cx, cy, cz;  // x0 < C-1, y0 < R-1, z0 < S-1 (or z0 < S when slices == 1: z1 = z0, cc:510-513)
    __device__ __forceinline__ explicit InsideK(const Geo &g)
        : cx(static_cast<unsigned>(g.C - 1)), cy(static_cast<unsigned>(g.R - 1)),
          cz(static_cast<unsigned>(g.is2d ? g.Sg : g.Sg - 1)) {}  // GLOBAL slice index
    template <typename CT>
    __device__ __forceinline__ bool operator()(const Tri<CT> &t) const {
        return t.ok && static_cast<unsigned>(t.x0) < cx && static_cast<unsigned>(t.y0) < cy &&
               static_cast<unsigned>(t.z0) < cz;
    }
};

// The eight corners of a sample point whose corners are ALL inside the grid: one row offset, three adds, and the
// x+1 neighbours as immediate offsets.  No clamps, no predicates.
template <typename IX, bool CH1>
__device__ __forceinline__ void load_corners(const float *__restrict__ M, const Geo &g, int x0, int y0, int z0, int c,
                                             float (&v)[8]) {
    const IX o00 = (static_cast<IX>(z0) * g.R + y0) * g.C + x0;
    const IX zstep = g.is2d ? IX(0) : static_cast<IX>(g.R) * g.C;
    if (CH1) {
        const float *p00 = M + o00;
        const float *p10 = p00 + g.C;
        const float *p01 = p00 + zstep;
        const float *p11 = p01 + g.C;
        v[0] = __ldg(p00);
        v[1] = __ldg(p00 + 1);
        v[2] = __ldg(p10);
        v[3] = __ldg(p10 + 1);
        v[4] = __ldg(p01);
        v[5] = __ldg(p01 + 1);
        v[6] = __ldg(p11);
        v[7] = __ldg(p11 + 1);
    } else {
        const IX ch = g.CH;
        const float *p00 = M + o00 * ch + c;
        const float *p10 = p00 + g.C * ch;
        const float *p01 = p00 + zstep * ch;
        const float *p11 = p01 + g.C * ch;
        v[0] = __ldg(p00);
        v[1] = __ldg(p00 + ch);
        v[2] = __ldg(p10);
        v[3] = __ldg(p10 + ch);
        v[4] = __ldg(p01);
        v[5] = __ldg(p01 + ch);
        v[6] = __ldg(p11);
        v[7] = __ldg(p11 + ch);
    }
}

// v = {c000, c100, c010, c110, c001, c101, c011, c111}
__device__ __forceinline__ float warp_interp(const float (&v)[8], const Tri<double> &t) {
    const double tx = t.tx, ty = t.ty, tz = t.tz;
    const double c00 = static_cast<double>(v[0]) * (1.0 - tx) + static_cast<double>(v[1]) * tx;
    const double c10 = static_cast<double>(v[2]) * (1.0 - tx) + static_cast<double>(v[3]) * tx;
    const double c01 = static_cast<double>(v[4]) * (1.0 - tx) + static_cast<double>(v[5]) * tx;
    const double c11 = static_cast<double>(v[6]) * (1.0 - tx) + static_cast<double>(v[7]) * tx;
    const double c0 = c00 * (1.0 - ty) + c10 * ty;
    const double c1 = c01 * (1.0 - ty) + c11 * ty;
    const float w = static_cast<float>(c0 * (1.0 - tz) + c1 * tz);
    return finite_f(w) ? w : CUDART_NAN_F;
}
__device__ __forceinline__ float warp_interp(const float (&v)[8], const Tri<float> &t) {
    const float b = v[0];
    const float tx = t.tx, ty = t.ty, tz = t.tz;
    const float e00 = tx * (v[1] - b);  // c000 - b == 0
    const float e10 = fmaf(tx, v[3] - v[2], v[2] - b);
    const float e01 = fmaf(tx, v[5] - v[4], v[4] - b);
    const float e11 = fmaf(tx, v[7] - v[6], v[6] - b);
    const float e0 = fmaf(ty, e10 - e00, e00);
    const float e1 = fmaf(ty, e11 - e01, e01);
    const float w = b + fmaf(tz, e1 - e0, e0);
    return finite_f(w) ? w : CUDART_NAN_F;
}

// `inside` is InsideK(t): a sample point with any corner outside the grid is NaN without touching memory (cc:516-522).
template <typename CT, typename IX, bool CH1>
__device__ __forceinline__ float warp_one(const float *__restrict__ M, const Geo &g, const Tri<CT> &t, bool inside, int c) {
    float w = CUDART_NAN_F;
    if (inside) {
        float v[8];
        load_corners<IX, CH1>(M, g, t.x0, t.y0, t.z0, c, v);
        w = warp_interp(v, t);
    }
    return w;
}

// ------------------------------------------------------------------------------------------------------------------
// compute_gradient for one voxel (channel 0 of the fixed image), cc:232-265: central difference in index units,
// one-sided (/1) at the borders, zero if a neighbour is non-finite or the extent is 1.  The per-thread neighbour
// offsets/scales along x and y are loop invariants of the caller.
// ------------------------------------------------------------------------------------------------------------------
template <typename IX>
struct GradK {
    IX xl, xr, yu, yd;      // element offsets of the in-plane neighbours (0 at the borders)
    float sx, sy;           // 1/(index distance): 0.5, 1 or 0 (extent 1)
    __device__ __forceinline__ GradK(const Geo &g, int x, int y, IX ch) {
        const int ixl = (x > 0) ? -1 : 0, ixr = (x + 1 < g.C) ? 1 : 0;
        const int iyu = (y > 0) ? -1 : 0, iyd = (y + 1 < g.R) ? 1 : 0;
        xl = ixl * ch;
        xr = ixr * ch;
        yu = static_cast<IX>(iyu) * g.C * ch;
        yd = static_cast<IX>(iyd) * g.C * ch;
        sx = (g.C > 1 && ixr != ixl) ? ((ixr - ixl == 2) ? 0.5f : 1.0f) : 0.0f;
        sy = (g.R > 1 && iyd != iyu) ? ((iyd - iyu == 2) ? 0.5f : 1.0f) : 0.0f;
    }
};

struct Stencil {  // the six neighbours of the fixed image, channel 0
    float xl, xr, yu, yd, zp, zn;
    float sz;  // 1/(index distance) along z: 0.5, 1 or 0
};
// in-plane neighbours (L1/L2 hits); the z neighbours are supplied by the caller, which marches along z and keeps the
// previous / next slice values of its voxel column in registers
template <typename IX>
__device__ __forceinline__ void load_stencil_xy(Stencil &s, const float *__restrict__ Fv, const GradK<IX> &k) {
    s.xl = __ldg(Fv + k.xl);
    s.xr = __ldg(Fv + k.xr);
    s.yu = __ldg(Fv + k.yu);
    s.yd = __ldg(Fv + k.yd);
}
__device__ __forceinline__ float z_scale(const Geo &g, int zg /* global slice index */) {
    if (g.Sg <= 1) return 0.0f;
    return (zg > 0 && zg + 1 < g.Sg) ? 0.5f : 1.0f;  // Sg > 1: at a border the one-sided difference has distance 1
}
__device__ __forceinline__ void grad_fixed(const Stencil &s, float sx, float sy, double &gx, double &gy, double &gz) {
    // (vr - vl) / (xr - xl) with xr - xl in {1, 2}: multiplying by 1 or 0.5 is the same operation bit for bit;
    // scale 0 encodes "extent 1" (cc:233, 244, 255)
    gx = (sx != 0.0f) ? diff_exact(s.xr, s.xl) * static_cast<double>(sx) : 0.0;
    gy = (sy != 0.0f) ? diff_exact(s.yd, s.yu) * static_cast<double>(sy) : 0.0;
    gz = (s.sz != 0.0f) ? diff_exact(s.zn, s.zp) * static_cast<double>(s.sz) : 0.0;
}
__device__ __forceinline__ void grad_fixed(const Stencil &s, float sx, float sy, float &gx, float &gy, float &gz) {
    const float dx = s.xr - s.xl, dy = s.yd - s.yu, dz = s.zn - s.zp;
    gx = finite_f(dx) ? dx * sx : 0.0f;
    gy = finite_f(dy) ? dy * sy : 0.0f;
    gz = finite_f(dz) ? dz * s.sz : 0.0f;
}

// ------------------------------------------------------------------------------------------------------------------
// Demons force of one voxel, cc:297-330.
// ------------------------------------------------------------------------------------------------------------------
struct ForceK {
    double norm_eps;      // normalization_factor + 1e-10 (the same double add as cc:315)
    double inv_norm_eps;  // 1 / norm_eps
    double max_update;    // cc:281
    double mu2_lo;        // max_update^2 * (1 - 2^-50): below it sqrt(|u|^2) <= max_update for certain
};

template <bool STRICT>
__device__ __forceinline__ void demons_force(float fixed_val, float moving_val, double gx, double gy, double gz,
                                             const ForceK &k, double &ux, double &uy, double &uz, double &term,
                                             int &valid) {
    constexpr double epsilon = 1.0e-10;
    ux = uy = uz = 0.0;
    if (!(finite_f(fixed_val) && finite_f(moving_val))) {
        term = 0.0;
        valid = 0;
        return;
    }
    const double diff = diff_exact(fixed_val, moving_val);  // == (double)fixed - (double)moving
    term = diff * diff;
    valid = 1;
    const double gmag_sq = gx * gx + gy * gy + gz * gz;
    const double denom = gmag_sq + (STRICT ? (diff * diff) / k.norm_eps : (diff * diff) * k.inv_norm_eps);
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
        const double m2 = ux * ux + uy * uy + uz * uz;
        if (m2 > k.mu2_lo) {  // only then can sqrt(m2) exceed max_update (cc:320-326)
            const double mag = sqrt(m2);
            if (mag > k.max_update) {
                const double scale = k.max_update / mag;
                ux *= scale;
                uy *= scale;
                uz *= scale;
            }
        }
    }
}
template <bool STRICT>
__device__ __forceinline__ void demons_force(float fixed_val, float moving_val, float gx, float gy, float gz,
                                             const ForceK &k, float &ux, float &uy, float &uz, float &term,
                                             int &valid) {
    ux = uy = uz = 0.0f;
    if (!(finite_f(fixed_val) && finite_f(moving_val))) {
        term = 0.0f;
        valid = 0;
        return;
    }
    const float diff = fixed_val - moving_val;
    term = diff * diff;
    valid = 1;
    const float denom = fmaf(gx, gx, fmaf(gy, gy, gz * gz)) + term * static_cast<float>(k.inv_norm_eps);
    if (denom > 1.0e-10f) {
        const float s = diff / denom;
        ux = s * gx;
        uy = s * gy;
        uz = s * gz;
        const float m2 = fmaf(ux, ux, fmaf(uy, uy, uz * uz));
        const float mu = static_cast<float>(k.max_update);
        if (m2 > mu * mu) {
            const float scale = mu * rsqrtf(m2);
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
// 3-D grid striding over slices (no div/mod); 32-bit element offsets (IX = int) whenever N*channels < 2^31; the fixed
// gradient is recomputed from F (7-point stencil served by L1/L2) instead of being read back as 3 field components.
// MSE terms: per-thread sum -> warp shuffle -> block -> one (sum, count) partial per block (deterministic).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void reduce_final_block(const int tid, const double *part_sum, const long long *part_cnt,
                                                   const int nparts, Ctrl *ctrl, double *mse_hist, const long long it,
                                                   const long long hist_idx, const int count_iter);

// MSE of one FORCE stage from (sum of squared differences, valid count) and the loop's stop rule (cc:340-342, 987-994).
//   count_iter != 0: this FORCE stage belongs to loop iteration `it` (iters_done, history and convergence updated)
__device__ __forceinline__ void apply_stop_rule(Ctrl *ctrl, double sum, long long n, double *mse_hist, const long long it,
                                                const long long hist_idx, const int count_iter) {
    double mse = sum;
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

#ifndef DCMA_FORCE_MINB64
#define DCMA_FORCE_MINB64 4  // fp64: 64 registers, 4 blocks/SM measured 16 % faster than 80 registers, 3 blocks/SM
#endif
// TU: storage type of the update field (DCMA_B200_FIELD_MIXED: T = double for D and all arithmetic, TU = float)
template <typename T, bool STRICT, typename IX, bool CH1, bool SYM, typename TU = T>
__global__ void __launch_bounds__(kTileX *kTileY, (sizeof(T) == 8 ? DCMA_FORCE_MINB64 : 4))
    k_warp_force(const float *__restrict__ F, const float *__restrict__ M, const float *__restrict__ Wbuf,
                 const T *__restrict__ D, TU *__restrict__ U, const Geo g, const int zlen, const int w_src,
                 const ForceK fk, double *part_sum, long long *part_cnt, Ctrl *ctrl, const long long it,
                 const int fused_reduce, double *mse_hist, const long long hist_idx, const int count_iter) {
    constexpr bool symmetric = SYM;  // extension a11: J = (grad F + grad W)/2; needs w_src == 2
    using CT = T;  // compute type of coordinates / interpolation / force == field storage type
    if (it > ctrl->converged_at) return;
    CT acc = 0;
    int cnt = 0;
    const int x = blockIdx.x * kTileX + threadIdx.x;
    const int y = blockIdx.y * kTileY + threadIdx.y;
    if (x < g.C && y < g.R) {
        const IX ch = CH1 ? IX(1) : static_cast<IX>(g.CH);
        const IX plane = static_cast<IX>(g.R) * g.C;
        const IX vxy = static_cast<IX>(y) * g.C + x;
        const T *D1 = D + g.N, *D2 = D + 2 * g.N;
        TU *U1 = U + g.N, *U2 = U + 2 * g.N;
        const AxisK<CT> ak(g);
        const InsideK ik(g);
        const GradK<IX> gk(g, x, y, ch);
        const CT xd = static_cast<CT>(x), yd = static_cast<CT>(y);
        // Each block marches through a run of CONSECUTIVE slices [z_begin, z_end): the z neighbours of the gradient
        // stencil roll through registers (one new F load per voxel instead of three), the moving-image gathers of
        // consecutive slices overlap in L1/L2, and the displacement of the next slice is requested one step ahead.
        const int z_begin = g.zb + blockIdx.z * zlen, z_end = min(z_begin + zlen, g.ze);
        if (z_begin < z_end) {
            IX v = static_cast<IX>(z_begin) * plane + vxy;
            const IX pch = plane * ch;
            CT zd = static_cast<CT>(z_begin + g.zoff);  // sample coordinates and the moving image are GLOBAL
            CT dx = 0, dy = 0, dz = 0;
            if (w_src == 1) {
                dx = D[v];
                dy = D1[v];
                dz = D2[v];
            }
            float f_prev = __ldg(F + (z_begin + g.zoff > 0 ? v - plane : v) * ch);
            float f_cur = __ldg(F + v * ch);
            float w_prev = 0.f, w_cur = 0.f;  // rolling z stencil of W (symmetric force only)
            if (symmetric) {
                w_prev = __ldg(Wbuf + (z_begin + g.zoff > 0 ? v - plane : v) * ch);
                w_cur = __ldg(Wbuf + v * ch);
            }
            // L2 prefetch, one instruction per warp and step: a tile row is one warp, and lane l asks for ONE line of
            // the row kPrefetchAhead slices ahead -- lane 0: F, lane 1: M (or W), lanes 2..: the lines of D's components.
            constexpr int NL = (sizeof(T) * kTileX + 127) / 128;  // 128-byte lines per field row segment
            const char *pf_ptr = nullptr;
            long long pf_step = 0;
            {
                const int lane = threadIdx.x;
                const long long row = (static_cast<long long>(z_begin) + kPrefetchAhead) * plane +
                                      static_cast<long long>(y) * g.C + blockIdx.x * kTileX;
                if (lane < 2) {
                    const float *base = (lane == 0) ? F : ((w_src == 2) ? Wbuf : M + static_cast<long long>(g.zoff) * plane * ch);
                    pf_ptr = reinterpret_cast<const char *>(base + row * ch);
                    pf_step = static_cast<long long>(plane) * ch * static_cast<long long>(sizeof(float));
                } else if (w_src == 1 && lane < 2 + 3 * NL) {
                    const int c = (lane - 2) / NL, h = (lane - 2) - c * NL;
                    pf_ptr = reinterpret_cast<const char *>(D + c * g.N + row) + h * 128;
                    pf_step = static_cast<long long>(plane) * static_cast<long long>(sizeof(T));
                }
            }
            for (int z = z_begin; z < z_end; ++z, v += plane, zd += CT(1)) {
                const float *Fv = F + v * ch;
                if (kPrefetchAhead > 0) {
                    if (pf_ptr != nullptr && z + kPrefetchAhead < g.S) prefetch_line(pf_ptr);
                    pf_ptr += pf_step;
                }
                Stencil st;
                load_stencil_xy<IX>(st, Fv, gk);
                const float f_next = (z + g.zoff + 1 < g.Sg) ? __ldg(Fv + pch) : f_cur;
                st.zp = f_prev;
                st.zn = f_next;
                st.sz = z_scale(g, z + g.zoff);
                const float fixed_val = f_cur;
                float moving_val;
                if (w_src == 1) {
                    // a non-finite displacement component makes the voxel NaN (cc:494-500): it fails tri_coords'
                    // range guard; with slices == 1 the unused z component is folded into x for that purpose
                    if (g.is2d) dx += dz - dz;
                    const Tri<CT> t = tri_coords<CT, STRICT>(g, ak, x, y, z + g.zoff, xd, yd, zd, dx, dy, dz);
                    moving_val = warp_one<CT, IX, CH1>(M, g, t, ik(t), 0);  // only channel 0 drives the force (cc:294-295)
                    if (z + 1 < z_end) {
                        dx = D[v + plane];
                        dy = D1[v + plane];
                        dz = D2[v + plane];
                    }
                } else {
                    // W := M at the voxel itself (first iteration): M is the whole volume, W a local buffer
                    moving_val = __ldg((w_src == 0 ? M + static_cast<IX>(g.zoff) * pch : Wbuf) + v * ch);
                }
                CT gx, gy, gz;
                grad_fixed(st, gk.sx, gk.sy, gx, gy, gz);
                if (symmetric) {
                    // gradient of the warped image by the same rule (central difference in index units, one-sided at
                    // the borders, zero next to a non-finite voxel), then the mean of the two gradients
                    Stencil sw;
                    load_stencil_xy<IX>(sw, Wbuf + v * ch, gk);
                    const float w_next = (z + g.zoff + 1 < g.Sg) ? __ldg(Wbuf + v * ch + pch) : w_cur;
                    sw.zp = w_prev;
                    sw.zn = w_next;
                    sw.sz = st.sz;
                    CT hx, hy, hz;
                    grad_fixed(sw, gk.sx, gk.sy, hx, hy, hz);
                    gx = (gx + hx) * CT(0.5);
                    gy = (gy + hy) * CT(0.5);
                    gz = (gz + hz) * CT(0.5);
                    w_prev = w_cur;
                    w_cur = w_next;
                }
                CT ux, uy, uz, term;
                int valid;
                demons_force<STRICT>(fixed_val, moving_val, gx, gy, gz, fk, ux, uy, uz, term, valid);
                U[v] = static_cast<TU>(ux);
                U1[v] = static_cast<TU>(uy);
                U2[v] = static_cast<TU>(uz);
                acc += term;
                cnt += valid;
                f_prev = f_cur;
                f_cur = f_next;
            }
        }
    }
    // warp-shuffle + block reduction of (sum of squared differences, valid count): cc:334-339 done on the device
    double a = static_cast<double>(acc);
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    __shared__ double s_acc[kTileY];
    __shared__ int s_cnt[kTileY];
    if (threadIdx.x == 0) {
        s_acc[threadIdx.y] = a;
        s_cnt[threadIdx.y] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        double aa = 0.0;
        long long c = 0;
        for (int i = 0; i < kTileY; ++i) {
            aa += s_acc[i];
            c += s_cnt[i];
        }
        const int b = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        part_sum[b] = aa;
        part_cnt[b] = c;
    }
    // Fused final reduction: the LAST block to publish its partial reduces all of them (in the fixed order of
    // k_reduce_final, so the result does not depend on which block that is) and applies the stop rule.
    if (fused_reduce) {
        __shared__ int s_last;
        const int nblocks = gridDim.x * gridDim.y * gridDim.z;
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            // RELEASE add: orders this block's partial before its ticket with a MEMBAR only.  (__threadfence() would
            // also invalidate the SM's L1 -- CCTL.IVALL -- once per block, throwing away the gather working set of the
            // blocks still running there; measured +6 % on the kernel.)  The last block alone pays the acquire side.
            unsigned t;
            asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(&ctrl->ticket) : "memory");
            s_last = (t == static_cast<unsigned>(nblocks - 1));
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (threadIdx.x == 0 && threadIdx.y == 0) ctrl->ticket = 0;
            reduce_final_block(threadIdx.y * kTileX + threadIdx.x, part_sum, part_cnt, nblocks, ctrl, mse_hist, it,
                               hist_idx, count_iter);
        }
    }
}

// Final reduction + the loop's stop rule (cc:340-342 and cc:987-994) by ONE 256-thread block, fixed order =>
// deterministic.  `tid` = linear thread index in the block.  Called by every thread of the block.
//   count_iter != 0: this FORCE stage belongs to loop iteration `it` (iters_done, history and convergence updated)
__device__ __forceinline__ void reduce_final_block(const int tid, const double *part_sum, const long long *part_cnt,
                                                   const int nparts, Ctrl *ctrl, double *mse_hist, const long long it,
                                                   const long long hist_idx, const int count_iter) {
    __shared__ double s_acc[256];
    __shared__ long long s_cnt[256];
    double a = 0.0;
    long long c = 0;
    for (int i = tid; i < nparts; i += 256) {
        a += part_sum[i];
        c += part_cnt[i];
    }
    s_acc[tid] = a;
    s_cnt[tid] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
            s_acc[tid] += s_acc[tid + o];
            s_cnt[tid] += s_cnt[tid + o];
        }
        __syncthreads();
    }
    if (tid == 0) {
        ctrl->last_sum = s_acc[0];
        apply_stop_rule(ctrl, s_acc[0], s_cnt[0], mse_hist, it, hist_idx, count_iter);
    }
}

// z-slab runs: every slab publishes (sum of squared differences, valid count) of its owned slices; after they have been
// gathered (NCCL all-gather or peer copies) each engine applies the stop rule to the totals, summed in rank order so
// that all ranks take the same decision bit for bit.   gather = world x {sum, (double)count}.
__global__ void k_stop_rule(Ctrl *__restrict__ ctrl, const double *__restrict__ gather, const int world,
                            double *__restrict__ mse_hist, const long long it, const long long hist_idx) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (it > ctrl->converged_at) return;
    double sum = 0.0, cnt = 0.0;
    for (int r = 0; r < world; ++r) {
        sum += gather[2 * r];
        cnt += gather[2 * r + 1];
    }
    apply_stop_rule(ctrl, sum, static_cast<long long>(cnt), mse_hist, it, hist_idx, 1);
}

// this slab's (sum, count) of the most recent FORCE stage, as two doubles (the element type of the gather)
__global__ void k_publish_mse(const Ctrl *__restrict__ ctrl, double *__restrict__ out2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    out2[0] = ctrl->last_sum;
    out2[1] = static_cast<double>(ctrl->last_n);
}

// the same as its own launch (stage API, unfused mode)
__global__ void __launch_bounds__(256)
    k_reduce_final(const double *__restrict__ part_sum, const long long *__restrict__ part_cnt, const int nparts,
                   Ctrl *__restrict__ ctrl, double *__restrict__ mse_hist, const long long it, const long long hist_idx,
                   const int count_iter) {
    if (it > ctrl->converged_at) return;
    reduce_final_block(threadIdx.x, part_sum, part_cnt, nparts, ctrl, mse_hist, it, hist_idx, count_iter);
}

// ------------------------------------------------------------------------------------------------------------------
// k_warp: warp_moving as its own kernel (stage API, KeepDeformedImages, dcma_b200_warp_image): all channels.
// ------------------------------------------------------------------------------------------------------------------
template <typename T, bool STRICT, typename IX, bool CH1>
__global__ void __launch_bounds__(kTileX *kTileY)
    k_warp(const float *__restrict__ M, const T *__restrict__ D, float *__restrict__ W, const Geo g, const int zlen,
           const int use_oob, const float oob_value) {
    using CT = T;
    const int x = blockIdx.x * kTileX + threadIdx.x;
    const int y = blockIdx.y * kTileY + threadIdx.y;
    if (x >= g.C || y >= g.R) return;
    const IX ch = CH1 ? IX(1) : static_cast<IX>(g.CH);
    const IX plane = static_cast<IX>(g.R) * g.C;
    const IX vxy = static_cast<IX>(y) * g.C + x;
    const T *D1 = D + g.N, *D2 = D + 2 * g.N;
    const AxisK<CT> ak(g);
    const InsideK ik(g);
    const CT xd = static_cast<CT>(x), yd = static_cast<CT>(y);
    const int z_begin = g.zb + blockIdx.z * zlen, z_end = min(z_begin + zlen, g.ze);
    for (int z = z_begin; z < z_end; ++z) {
        const IX v = static_cast<IX>(z) * plane + vxy;
        CT dx = D[v];
        const CT dy = D1[v], dz = D2[v];
        if (g.is2d) dx += dz - dz;  // non-finite z component => NaN voxel even when z is not interpolated (cc:494-500)
        const Tri<CT> t = tri_coords<CT, STRICT>(g, ak, x, y, z + g.zoff, xd, yd, static_cast<CT>(z + g.zoff), dx, dy, dz);
        const bool inside = ik(t);
        const int nch = CH1 ? 1 : g.CH;
        for (int c = 0; c < nch; ++c) {
            float w = warp_one<CT, IX, CH1>(M, g, t, inside, c);
            if (use_oob && !finite_f(w)) w = oob_value;
            W[v * ch + c] = w;
        }
    }
}

// k_gradient: compute_gradient as its own kernel, AoS double output (stage API / inspection only).
__global__ void __launch_bounds__(kTileX *kTileY)
    k_gradient(const float *__restrict__ F, double *__restrict__ G_aos, const Geo g, const int zlen) {
    const int x = blockIdx.x * kTileX + threadIdx.x;
    const int y = blockIdx.y * kTileY + threadIdx.y;
    if (x >= g.C || y >= g.R) return;
    const long long ch = g.CH;
    const long long plane = static_cast<long long>(g.R) * g.C;
    const GradK<long long> gk(g, x, y, ch);
    const int z_begin = g.zb + blockIdx.z * zlen, z_end = min(z_begin + zlen, g.ze);
    for (int z = z_begin; z < z_end; ++z) {
        const long long v = static_cast<long long>(z) * plane + static_cast<long long>(y) * g.C + x;
        const float *Fv = F + v * ch;
        Stencil st;
        load_stencil_xy<long long>(st, Fv, gk);
        st.zp = __ldg(Fv - ((z + g.zoff > 0) ? plane * ch : 0));
        st.zn = __ldg(Fv + ((z + g.zoff + 1 < g.Sg) ? plane * ch : 0));
        st.sz = z_scale(g, z + g.zoff);
        double gx, gy, gz;
        grad_fixed(st, gk.sx, gk.sy, gx, gy, gz);
        G_aos[v * 3 + 0] = gx;
        G_aos[v * 3 + 1] = gy;
        G_aos[v * 3 + 2] = gz;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_add_update: D[c] += U[c] * pxl_c   (cc:361-366)
// ------------------------------------------------------------------------------------------------------------------
template <typename T, typename TU = T>
__global__ void __launch_bounds__(256)
    k_add_update(T *__restrict__ D, const TU *__restrict__ U, const long long N, const double px, const double py,
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
// no NaN checks; the four upper-z corners are narrowed to float as the reference does (cc:439-442).
// Reads D only at self (in place is race-free), gathers U at 8 corners x 3 components; the corner offsets and the
// in-grid predicates are computed once and shared by the three components.
// ------------------------------------------------------------------------------------------------------------------
// fp64: 24 gathered doubles per voxel put the kernel at ~100 registers (2 blocks/SM); capping it at 85 (3 blocks/SM)
// measured 3 % faster.  fp32 needs no cap (48 registers).
template <typename T, bool STRICT, typename IX, typename TU = T>
__global__ void __launch_bounds__(kTileX *kTileY, (sizeof(T) == 8 ? 3 : 4))
    k_compose(T *__restrict__ D, const TU *__restrict__ U, const Geo g, const int zlen, Ctrl *__restrict__ ctrl,
              const long long it) {
    using CT = T;
    if (it > ctrl->converged_at) return;
    const int x = blockIdx.x * kTileX + threadIdx.x;
    const int y = blockIdx.y * kTileY + threadIdx.y;
    if (x >= g.C || y >= g.R) return;
    const IX plane = static_cast<IX>(g.R) * g.C;
    const IX vxy = static_cast<IX>(y) * g.C + x;
    const AxisK<CT> ak(g);
    const InsideK ik(g);
    const CT xd = static_cast<CT>(x), yd = static_cast<CT>(y);
    const long long N = g.N;
    const int z_begin = g.zb + blockIdx.z * zlen, z_end = min(z_begin + zlen, g.ze);  // consecutive slices: gathers overlap
    if (z_begin >= z_end) return;
    CT zd = static_cast<CT>(z_begin + g.zoff);  // sample coordinates are GLOBAL; U and D are indexed locally
    CT d[3];
    {
        const IX v0 = static_cast<IX>(z_begin) * plane + vxy;
#pragma unroll
        for (int c = 0; c < 3; ++c) d[c] = D[v0 + c * N];
    }
    const CT pxl[3] = {ak.px, ak.py, ak.pz};
    // L2 prefetch, one instruction per warp and step (see k_warp_force): lanes 0..3*NL-1 own the lines of U's three
    // components, the next 3*NL lanes those of D's, kPrefetchAhead slices ahead of the slice being composed.
    constexpr int NL = (sizeof(T) * kTileX + 127) / 128, NLU = (sizeof(TU) * kTileX + 127) / 128;
    static_assert(3 * (NL + NLU) <= kTileX, "one prefetch lane per line");
    const char *pf_ptr = nullptr;
    long long pf_step = 0;
    {
        const int lane = threadIdx.x;
        const long long row = (static_cast<long long>(z_begin) + kPrefetchAhead) * plane +
                              static_cast<long long>(y) * g.C + blockIdx.x * kTileX;
        if (lane < 3 * NLU) {
            const int c = lane / NLU, h = lane - c * NLU;
            pf_ptr = reinterpret_cast<const char *>(U + c * N + row) + h * 128;
            pf_step = static_cast<long long>(plane) * static_cast<long long>(sizeof(TU));
        } else if (lane < 3 * (NLU + NL)) {
            const int r = lane - 3 * NLU, c = r / NL, h = r - c * NL;
            pf_ptr = reinterpret_cast<const char *>(D + c * N + row) + h * 128;
            pf_step = static_cast<long long>(plane) * static_cast<long long>(sizeof(T));
        }
    }
    for (int z = z_begin; z < z_end; ++z, zd += CT(1)) {
        const IX v = static_cast<IX>(z) * plane + vxy;
        if (kPrefetchAhead > 0) {
            if (pf_ptr != nullptr && z + kPrefetchAhead < g.S) prefetch_line(pf_ptr);
            pf_ptr += pf_step;
        }
        const Tri<CT> t = tri_coords<CT, STRICT>(g, ak, x, y, z + g.zoff, xd, yd, zd, d[0], d[1], d[2]);
        const CT tx = t.tx, ty = t.ty, tz = t.tz;
        // DCMA_B200_FIELD_MIXED (U stored as float, D as double): the eight gathered values stay float and are interpolated
        // in float; one conversion per component instead of 24 (F2F.F64.F32 runs on the quarter-rate XU pipe: ncu showed
        // it 45 % busy with the conversions).  The reference itself narrows half of these values to float (cc:439-442).
        constexpr bool kNarrowU = sizeof(TU) < sizeof(T);
        using QT = typename std::conditional<kNarrowU, TU, CT>::type;
        QT q[3][8];
        const int lz0 = t.z0 - g.zoff;  // local slice of the lower corners
        // all corners inside the grid AND inside this slab's buffers (always true for a whole-volume engine)
        const bool inside = ik(t) && static_cast<unsigned>(lz0) < static_cast<unsigned>(g.is2d ? g.S : g.S - 1);
        if (inside) {
            // all eight corners in the grid (the common case): one row offset shared by the three components,
            // x+1 neighbours as immediate offsets, no predicates
            const IX o00 = (static_cast<IX>(lz0) * g.R + t.y0) * g.C + t.x0;
            const IX zstep = g.is2d ? IX(0) : plane;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const TU *p00 = U + c * N + o00;
                const TU *p10 = p00 + g.C;
                const TU *p01 = p00 + zstep;
                const TU *p11 = p01 + g.C;
                q[c][0] = p00[0];
                q[c][1] = p00[1];
                q[c][2] = p10[0];
                q[c][3] = p10[1];
                q[c][4] = p01[0];
                q[c][5] = p01[1];
                q[c][6] = p11[0];
                q[c][7] = p11[1];
            }
        } else {
            // some corner outside: gather from clamped coordinates, then replace outside corners by zero (cc:424-427).
            // A coordinate that failed the range guard has every corner outside.
            const int x1 = t.x0 + 1, y1 = t.y0 + 1, z1 = g.is2d ? t.z0 : (t.z0 + 1);
            const bool vx0 = t.ok && static_cast<unsigned>(t.x0) < static_cast<unsigned>(g.C);
            const bool vx1 = t.ok && static_cast<unsigned>(x1) < static_cast<unsigned>(g.C);
            const bool vy0 = static_cast<unsigned>(t.y0) < static_cast<unsigned>(g.R);
            const bool vy1 = static_cast<unsigned>(y1) < static_cast<unsigned>(g.R);
            const bool vz0 = static_cast<unsigned>(t.z0) < static_cast<unsigned>(g.Sg);
            const bool vz1 = static_cast<unsigned>(z1) < static_cast<unsigned>(g.Sg);
            const int lz1 = z1 - g.zoff;
            // a corner inside the grid whose plane this slab does not hold: the halo is too thin for this displacement
            if (t.ok && ((vz0 && static_cast<unsigned>(lz0) >= static_cast<unsigned>(g.S)) ||
                         (vz1 && static_cast<unsigned>(lz1) >= static_cast<unsigned>(g.S))))
                atomicOr(&ctrl->halo_overflow, 1u);
            const int cx0 = min(max(t.x0, 0), g.C - 1), cx1 = min(max(x1, 0), g.C - 1);
            const int cy0 = min(max(t.y0, 0), g.R - 1), cy1 = min(max(y1, 0), g.R - 1);
            const int cz0 = min(max(lz0, 0), g.S - 1), cz1 = min(max(lz1, 0), g.S - 1);
            const IX r00 = (static_cast<IX>(cz0) * g.R + cy0) * g.C, r10 = (static_cast<IX>(cz0) * g.R + cy1) * g.C;
            const IX r01 = (static_cast<IX>(cz1) * g.R + cy0) * g.C, r11 = (static_cast<IX>(cz1) * g.R + cy1) * g.C;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const TU *Uc = U + c * N;
                q[c][0] = (vz0 && vy0 && vx0) ? static_cast<QT>(Uc[r00 + cx0]) : QT(0);
                q[c][1] = (vz0 && vy0 && vx1) ? static_cast<QT>(Uc[r00 + cx1]) : QT(0);
                q[c][2] = (vz0 && vy1 && vx0) ? static_cast<QT>(Uc[r10 + cx0]) : QT(0);
                q[c][3] = (vz0 && vy1 && vx1) ? static_cast<QT>(Uc[r10 + cx1]) : QT(0);
                // slices == 1: z1 == z0, the z0 plane is reused (cc:436-442)
                q[c][4] = (vz1 && vy0 && vx0) ? static_cast<QT>(Uc[r01 + cx0]) : QT(0);
                q[c][5] = (vz1 && vy0 && vx1) ? static_cast<QT>(Uc[r01 + cx1]) : QT(0);
                q[c][6] = (vz1 && vy1 && vx0) ? static_cast<QT>(Uc[r11 + cx0]) : QT(0);
                q[c][7] = (vz1 && vy1 && vx1) ? static_cast<QT>(Uc[r11 + cx1]) : QT(0);
            }
        }
        // prefetch the next slice's displacement before consuming the gathers
        CT dn[3] = {CT(0), CT(0), CT(0)};
        if (z + 1 < z_end) {
            const IX vn = v + plane;
#pragma unroll
            for (int c = 0; c < 3; ++c) dn[c] = D[vn + c * N];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if constexpr (kNarrowU) {
                const QT fx = static_cast<QT>(tx), fy = static_cast<QT>(ty), fz = static_cast<QT>(tz);
                const QT c00 = q[c][0] * (QT(1) - fx) + q[c][1] * fx;
                const QT c10 = q[c][2] * (QT(1) - fx) + q[c][3] * fx;
                const QT c01 = q[c][4] * (QT(1) - fx) + q[c][5] * fx;
                const QT c11 = q[c][6] * (QT(1) - fx) + q[c][7] * fx;
                const QT c0 = c00 * (QT(1) - fy) + c10 * fy;
                const QT c1 = c01 * (QT(1) - fy) + c11 * fy;
                const CT u = static_cast<CT>(c0 * (QT(1) - fz) + c1 * fz);
                D[v + c * N] = d[c] + u * pxl[c];  // cc:460-462
                continue;
            }
            const CT c000 = q[c][0], c100 = q[c][1], c010 = q[c][2], c110 = q[c][3];
            // the four upper-z corners are narrowed to float as the reference does (cc:439-442)
            const CT c001 = narrow_to_float<STRICT>(q[c][4]);
            const CT c101 = narrow_to_float<STRICT>(q[c][5]);
            const CT c011 = narrow_to_float<STRICT>(q[c][6]);
            const CT c111 = narrow_to_float<STRICT>(q[c][7]);
            const CT c00 = c000 * (CT(1) - tx) + c100 * tx;  // cc:445-451
            const CT c10 = c010 * (CT(1) - tx) + c110 * tx;
            const CT c01 = c001 * (CT(1) - tx) + c101 * tx;
            const CT c11 = c011 * (CT(1) - tx) + c111 * tx;
            const CT c0 = c00 * (CT(1) - ty) + c10 * ty;
            const CT c1 = c01 * (CT(1) - ty) + c11 * ty;
            const CT u = c0 * (CT(1) - tz) + c1 * tz;
            D[v + c * N] = d[c] + u * pxl[c];  // cc:460-462
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) d[c] = dn[c];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// boundary conversions: reference AoS double [z][y][x][3]  <->  device SoA T [3][z][y][x]
// ------------------------------------------------------------------------------------------------------------------
// voxels [v0, v1) of the SoA field <-> a dense AoS chunk of (v1 - v0) * 3 doubles
template <typename T>
__global__ void __launch_bounds__(256) k_export_aos(const T *__restrict__ soa, double *__restrict__ aos, const long long N,
                                                     const long long v0, const long long v1) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long n = 3 * (v1 - v0);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long v = i / 3;
        const int c = static_cast<int>(i - v * 3);
        aos[i] = static_cast<double>(soa[v0 + v + c * N]);  // coalesced write, 3-way strided (cached) read
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_import_aos(const double *__restrict__ aos, T *__restrict__ soa, const long long N,
                                                     const long long v0, const long long v1) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long n = 3 * (v1 - v0);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long v = i / 3;
        const int c = static_cast<int>(i - v * 3);
        soa[v0 + v + c * N] = static_cast<T>(aos[i]);
    }
}

}  // namespace DCMA_NS
