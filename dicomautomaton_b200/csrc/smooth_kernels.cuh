// dcma_b200 -- separable 3-D Gaussian smoothing of a 3-component field: smooth_on_device,
// reference src/Alignment_Demons.cc:563-702.
//
// Reference semantics (per axis pass, cc:619-634 / 648-663 / 677-692):
//     out = sum_k w[k]*v[pos+k] / sum_k w[k]    over taps that are inside the grid AND finite, k = -r..r in order;
//     if no tap qualifies the input value is kept.   r = max(1, int(3*sigma/pxl)) per axis (cc:574-576).
// x pass, then y pass, then z pass, result back in the same buffer (cc:696).
//
// Two implementations with identical results:
//  * k_smooth_axis   -- one axis per launch, runtime radius, per-tap bounds+finite tests (the literal restatement;
//                       used for anisotropic/large radii and as the `fused=0` path).
//  * k_smooth3d      -- ONE launch for all three passes: a CTA owns an x-y tile, marches along z, stages each input
//                       slice tile (+halo) in shared memory with cp.async (double buffered), does the x and y passes
//                       in shared memory and keeps the last 2r+1 xy-smoothed values of its voxels in REGISTERS for the
//                       z pass.  Every field element is read from HBM once (+halo, served by L2) and written once.
//    Out-of-grid taps are staged as 0 and the per-position weight sums come from host tables built by in-order
//    accumulation, so for finite fields the arithmetic (order included) equals the reference's; a non-finite result
//    re-runs that output through the literal per-tap path, which makes the two implementations agree on any input.
#pragma once
#include "common.cuh"

#ifndef DCMA_NS
#error "define DCMA_NS (one namespace per precision-mode translation unit) before including"
#endif
namespace DCMA_NS {
using namespace ::dcma;

constexpr int kMaxFastRadius = 6;

// Per-launch constants of one smoothing call (kernel parameter space => uniform/constant-bank reads).
template <typename T>
struct SmoothWeights {
    T w[3][2 * kMaxFastRadius + 1];  // [axis: 0=x,1=y,2=z][k + r]
};

// STRICT: table holds wsum (divide, as the reference does); otherwise it holds 1/wsum (multiply).
template <typename T, bool STRICT>
__device__ __forceinline__ T normalise(T sum, T tab) {
    return STRICT ? sum / tab : sum * tab;
}

// ------------------------------------------------------------------------------------------------------------------
// literal single-axis pass (any radius): thread per element of the SoA field (3*N elements)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    k_smooth_axis(const T *__restrict__ in, T *__restrict__ out, const Geo g, const int axis, const int rad,
                  const double *__restrict__ w /* 2*rad+1 doubles in global memory */, const Ctrl *__restrict__ ctrl,
                  const long long it) {
    if (it > ctrl->converged_at) return;
    const long long N = g.N;
    const long long stride_thr = static_cast<long long>(gridDim.x) * blockDim.x;
    const int extent = (axis == 0) ? g.C : (axis == 1) ? g.R : g.S;
    const long long stride = (axis == 0) ? 1 : (axis == 1) ? g.C : static_cast<long long>(g.R) * g.C;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < 3 * N; i += stride_thr) {
        const long long v = i % N;
        const int x = static_cast<int>(v % g.C);
        const long long r1 = v / g.C;
        const int y = static_cast<int>(r1 % g.R);
        const int z = static_cast<int>(r1 / g.R);
        const int pos = (axis == 0) ? x : (axis == 1) ? y : z;
        T sum = 0, wsum = 0;
        for (int k = -rad; k <= rad; ++k) {
            const int q = pos + k;
            if (q >= 0 && q < extent) {
                const T val = in[i + k * stride];
                if (isfinite(val)) {
                    const T wk = static_cast<T>(w[k + rad]);
                    sum += wk * val;
                    wsum += wk;
                }
            }
        }
        out[i] = (wsum > T(0)) ? (sum / wsum) : in[i];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// fused three-pass kernel
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
struct alignas(4 * sizeof(T)) Vec4 {
    T v[4];
};

__device__ __forceinline__ void cp_async_zfill(void *smem_dst, const void *gmem_src, int bytes, int src_bytes) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    if (bytes == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
    else if (bytes == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

template <typename T, int R, int TX, int TY, int NT>
struct Smooth3dCfg {
    static constexpr int HX = ((R + 3) / 4) * 4;       // x halo rounded up so rows stay 16-byte aligned
    static constexpr int PW = TX + 2 * HX;             // staged row width
    static constexpr int PH = TY + 2 * R;              // staged rows
    static constexpr int QX = TX / 4;                  // quads (4 consecutive x) per tile row
    static constexpr int RG = NT / QX;                 // row groups
    static constexpr int OY = TY / RG;                 // rows owned per thread
    static constexpr int TAPS = 2 * R + 1;
    static constexpr size_t smem_bytes = (size_t(2) * PH * PW + size_t(PH) * TX) * sizeof(T);
    static_assert(TX % 4 == 0 && NT % QX == 0 && TY % RG == 0, "tile/thread shape");
};

// One output through the literal per-tap path (non-finite data only).  `tap(k)` returns the k-th tap value,
// `inb(k)` whether it lies inside the grid, `centre` the input value kept when no tap qualifies.
// Force-inlined and fully unrolled so the tap arrays stay in registers (a real call would force them to local memory).
template <typename T, int R, typename TapF, typename InbF>
__device__ __forceinline__ T smooth_slow(const T *w, TapF tap, InbF inb, T centre) {
    T sum = 0, wsum = 0;
#pragma unroll
    for (int k = 0; k < 2 * R + 1; ++k) {
        if (inb(k)) {
            const T val = tap(k);
            if (isfinite(val)) {
                sum += w[k] * val;
                wsum += w[k];
            }
        }
    }
    return (wsum > T(0)) ? (sum / wsum) : centre;
}

// grid: x = tiles along columns, y = tiles along rows, z = 3 components * z-chunks
template <typename T, bool STRICT, int R, int TX, int TY, int NT, bool VEC>
__global__ void __launch_bounds__(NT)
    k_smooth3d(const T *__restrict__ in, T *__restrict__ out, const Geo g, const SmoothWeights<T> W,
               const T *__restrict__ tab_x, const T *__restrict__ tab_y, const T *__restrict__ tab_z, const int zchunk,
               const int nzchunks, const Ctrl *__restrict__ ctrl, const long long it) {
    using Cfg = Smooth3dCfg<T, R, TX, TY, NT>;
    constexpr int HX = Cfg::HX, PW = Cfg::PW, PH = Cfg::PH, QX = Cfg::QX, RG = Cfg::RG, OY = Cfg::OY, TAPS = Cfg::TAPS;
    if (it > ctrl->converged_at) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *s_in = reinterpret_cast<T *>(smem_raw);   // [2][PH][PW]
    T *s_x = s_in + 2 * PH * PW;                 // [PH][TX]

    const int tid = threadIdx.x;
    const int comp = blockIdx.z / nzchunks;
    const int chunk = blockIdx.z - comp * nzchunks;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z_begin = chunk * zchunk;
    const int z_end = min(z_begin + zchunk, g.S);
    const long long plane = static_cast<long long>(g.R) * g.C;
    const T *src = in + comp * g.N;
    T *dst = out + comp * g.N;

    // stage one input slice tile (+halo) with cp.async; out-of-grid elements are zero-filled (src-size 0)
    auto stage = [&](int zp, int buf) {
        T *sb = s_in + buf * PH * PW;
        if (zp < 0 || zp >= g.S) return;  // nothing to load: the whole slice is outside (handled by the caller)
        const T *sp = src + zp * plane;
        if (VEC) {
            constexpr int E = 16 / sizeof(T);  // elements per 16-byte copy
            constexpr int CPR = PW / E;        // copies per staged row
            for (int e = tid; e < PH * CPR; e += NT) {
                const int row = e / CPR, cc = e - row * CPR;
                const int gy = y0 - R + row, gx = x0 - HX + cc * E;
                const bool ok = (gy >= 0) && (gy < g.R) && (gx >= 0) && (gx + E <= g.C);
                const T *p = ok ? (sp + static_cast<long long>(gy) * g.C + gx) : src;
                cp_async_zfill(sb + row * PW + cc * E, p, 16, ok ? 16 : 0);
            }
        } else {
            for (int e = tid; e < PH * PW; e += NT) {
                const int row = e / PW, col = e - row * PW;
                const int gy = y0 - R + row, gx = x0 - HX + col;
                const bool ok = (gy >= 0) && (gy < g.R) && (gx >= 0) && (gx < g.C);
                const T *p = ok ? (sp + static_cast<long long>(gy) * g.C + gx) : src;
                cp_async_zfill(sb + row * PW + col, p, static_cast<int>(sizeof(T)), ok ? static_cast<int>(sizeof(T)) : 0);
            }
        }
    };

    // z window: the last TAPS xy-smoothed values of each owned output, in registers
    T win[OY][4][TAPS];
#pragma unroll
    for (int j = 0; j < OY; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < TAPS; ++k) win[j][i][k] = T(0);

    const int qx = tid % QX, ry = tid / QX;
    const int gx_out = x0 + 4 * qx;

    const int zp_first = z_begin - R, zp_last = z_end - 1 + R;
    stage(zp_first, 0);
    cp_async_commit();

    for (int zp = zp_first; zp <= zp_last; ++zp) {
        const int buf = (zp - zp_first) & 1;
        cp_async_wait_all();
        __syncthreads();  // s_in[buf] complete; everyone is done with s_x and s_in[buf^1] of the previous step
        if (zp + 1 <= zp_last) stage(zp + 1, buf ^ 1);
        cp_async_commit();

        const bool slice_in = (zp >= 0 && zp < g.S);
        T vxy[OY][4];
        if (slice_in) {
            const T *sb = s_in + buf * PH * PW;
            // ---- x pass: s_in -> s_x, 4 consecutive outputs per task from HX/2+1 aligned 4-vectors
            for (int task = tid; task < PH * QX; task += NT) {
                const int row = task / QX, q = task - row * QX;
                constexpr int NV = (4 + 2 * HX) / 4;
                T a[4 * NV];
#pragma unroll
                for (int n = 0; n < NV; ++n) {
                    const Vec4<T> t4 = *reinterpret_cast<const Vec4<T> *>(sb + row * PW + 4 * q + 4 * n);
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[4 * n + i] = t4.v[i];
                }
                Vec4<T> o;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    T sum = 0;
#pragma unroll
                    for (int k = 0; k < TAPS; ++k) sum += W.w[0][k] * a[HX - R + i + k];
                    const int gx = x0 + 4 * q + i;
                    T res = normalise<T, STRICT>(sum, tab_x[min(gx, g.C - 1)]);
                    if (!isfinite(res)) {
                        res = smooth_slow<T, R>(
                            W.w[0], [&](int k) { return a[HX - R + i + k]; },
                            [&](int k) { const int q2 = gx + k - R; return q2 >= 0 && q2 < g.C; }, a[HX + i]);
                    }
                    o.v[i] = res;
                }
                *reinterpret_cast<Vec4<T> *>(s_x + row * TX + 4 * q) = o;
            }
            __syncthreads();
            // ---- y pass: s_x -> registers
#pragma unroll
            for (int j = 0; j < OY; ++j) {
                const int yl = ry + j * RG;  // local output row
                const int gy = y0 + yl;
                T col[TAPS][4];
#pragma unroll
                for (int k = 0; k < TAPS; ++k) {
                    const Vec4<T> t4 = *reinterpret_cast<const Vec4<T> *>(s_x + (yl + k) * TX + 4 * qx);
#pragma unroll
                    for (int i = 0; i < 4; ++i) col[k][i] = t4.v[i];
                }
                const T ty = tab_y[min(gy, g.R - 1)];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    T sum = 0;
#pragma unroll
                    for (int k = 0; k < TAPS; ++k) sum += W.w[1][k] * col[k][i];
                    T res = normalise<T, STRICT>(sum, ty);
                    if (!isfinite(res)) {
                        res = smooth_slow<T, R>(
                            W.w[1], [&](int k) { return col[k][i]; },
                            [&](int k) { const int q2 = gy + k - R; return q2 >= 0 && q2 < g.R; }, col[R][i]);
                    }
                    vxy[j][i] = res;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < OY; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) vxy[j][i] = T(0);
        }
        // ---- z pass from the register window
        const int zo = zp - R;  // output slice completed by this step
        const bool emit = (zo >= z_begin) && (zo < z_end);
        const T tz = tab_z[min(max(zo, 0), g.S - 1)];
#pragma unroll
        for (int j = 0; j < OY; ++j) {
            const int gy = y0 + ry + j * RG;
            Vec4<T> o;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int k = 0; k < TAPS - 1; ++k) win[j][i][k] = win[j][i][k + 1];
                win[j][i][TAPS - 1] = vxy[j][i];
                T sum = 0;
#pragma unroll
                for (int k = 0; k < TAPS; ++k) sum += W.w[2][k] * win[j][i][k];
                T res = normalise<T, STRICT>(sum, tz);
                if (!isfinite(res)) {
                    res = smooth_slow<T, R>(
                        W.w[2], [&](int k) { return win[j][i][k]; },
                        [&](int k) { const int q2 = zo + k - R; return q2 >= 0 && q2 < g.S; }, win[j][i][R]);
                }
                o.v[i] = res;
            }
            if (emit && gy < g.R) {
                T *p = dst + zo * plane + static_cast<long long>(gy) * g.C + gx_out;
                if (VEC) {
                    if (gx_out < g.C) *reinterpret_cast<Vec4<T> *>(p) = o;  // C % 4 == 0 in the VEC variant
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gx_out + i < g.C) p[i] = o.v[i];
                }
            }
        }
    }
}

}  // namespace DCMA_NS
