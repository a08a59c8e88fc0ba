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
#include <type_traits>

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
struct alignas(16) VecT {  // one 16-byte shared/global access: 4 floats or 2 doubles
    static constexpr int W = 16 / sizeof(T);
    T v[W];
};

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src, int src_bytes) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
template <int BYTES>
__device__ __forceinline__ void cp_async_small(void *smem_dst, const void *gmem_src, int src_bytes) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(s), "l"(gmem_src), "n"(BYTES), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

template <typename T, int R, int TX, int TY, int NT>
struct Smooth3dCfg {
    static constexpr int VW = 16 / sizeof(T);             // elements per 16-byte vector
    static constexpr int HX = ((R + VW - 1) / VW) * VW;   // x halo rounded up so staged rows stay 16-byte aligned
    static constexpr int PW = TX + 2 * HX;                // staged row width
    static constexpr int PH = TY + 2 * R;                 // staged rows
    static constexpr int QX = TX / VW;                    // vectors per tile row (thread column index = tid % QX)
    static constexpr int RG = NT / QX;                    // thread rows
    static constexpr int OY = TY / RG;                    // ADJACENT output rows owned by a thread
    static constexpr int TAPS = 2 * R + 1;
    static constexpr int XT = PH * QX;                    // x-pass tasks (one vector of outputs each)
    static constexpr int XN = (XT + NT - 1) / NT;
    static constexpr int NV = (VW + 2 * HX) / VW;         // vectors read per x-pass task
    static constexpr int CPR = PW / VW;                   // 16-byte copies per staged row
    static constexpr int NCP = (PH * CPR + NT - 1) / NT;  // 16-byte copies per thread per slice
    static constexpr size_t smem_bytes = (size_t(2) * PH * PW + size_t(PH) * TX) * sizeof(T);
    static_assert(TX % VW == 0 && NT % QX == 0 && TY % RG == 0 && (QX & (QX - 1)) == 0, "tile/thread shape");
};

// The literal definition for ONE output voxel, straight from global memory (cc:609-693): used only when the fast path
// produced a non-finite value, i.e. when some tap in the (2r+1)^3 support is non-finite.  Deliberately not inlined.
template <typename T, int R>
__device__ __noinline__ T smooth_literal(const T *__restrict__ src, const Geo g, const double *__restrict__ dwx,
                                         const double *__restrict__ dwy, const double *__restrict__ dwz, int x, int y,
                                         int z) {
    const long long plane = static_cast<long long>(g.R) * g.C;
    T zs = 0, zw = 0, z_centre = 0;
    for (int kz = -R; kz <= R; ++kz) {
        const int zz = z + kz;
        if (zz < 0 || zz >= g.S) continue;
        T ys = 0, yw = 0, y_centre = 0;
        for (int ky = -R; ky <= R; ++ky) {
            const int yy = y + ky;
            if (yy < 0 || yy >= g.R) continue;
            const T *row = src + zz * plane + static_cast<long long>(yy) * g.C;
            T xs = 0, xw = 0;
            for (int kx = -R; kx <= R; ++kx) {
                const int xx = x + kx;
                if (xx < 0 || xx >= g.C) continue;
                const T val = row[xx];
                if (isfinite(val)) {
                    const T wk = static_cast<T>(dwx[kx + R]);
                    xs += wk * val;
                    xw += wk;
                }
            }
            const T xv = (xw > T(0)) ? (xs / xw) : row[x];
            if (ky == 0) y_centre = xv;
            if (isfinite(xv)) {
                const T wk = static_cast<T>(dwy[ky + R]);
                ys += wk * xv;
                yw += wk;
            }
        }
        const T yv = (yw > T(0)) ? (ys / yw) : y_centre;
        if (kz == 0) z_centre = yv;
        if (isfinite(yv)) {
            const T wk = static_cast<T>(dwz[kz + R]);
            zs += wk * yv;
            zw += wk;
        }
    }
    return (zw > T(0)) ? (zs / zw) : z_centre;
}

// calls f(integral_constant<P>, zb + P) for P = 0..N-1 while zb + P <= last (compile-time ring phases)
template <int N, int P, typename F>
__device__ __forceinline__ void unroll_phases(F &f, const int zb, const int last) {
    if constexpr (P < N) {
        if (zb + P <= last) {
            f(std::integral_constant<int, P>{}, zb + P);
            unroll_phases<N, P + 1>(f, zb, last);
        }
    }
}

// grid: x = tiles along columns, y = tiles along rows, z = 3 components * z-chunks.
// tab_x / tab_y / tab_z are padded by the host to whole tiles (no bounds tests here).
template <typename T, bool STRICT, int R, int TX, int TY, int NT, int MINB, bool VEC>
__global__ void __launch_bounds__(NT, MINB)
    k_smooth3d(const T *__restrict__ in, T *__restrict__ out, const Geo g, const SmoothWeights<T> W,
               const T *__restrict__ tab_x, const T *__restrict__ tab_y, const T *__restrict__ tab_z,
               const double *__restrict__ dwx, const double *__restrict__ dwy, const double *__restrict__ dwz,
               const int zchunk, const int nzchunks, const Ctrl *__restrict__ ctrl, const long long it) {
    using Cfg = Smooth3dCfg<T, R, TX, TY, NT>;
    constexpr int VW = Cfg::VW, HX = Cfg::HX, PW = Cfg::PW, PH = Cfg::PH, QX = Cfg::QX, OY = Cfg::OY, TAPS = Cfg::TAPS;
    constexpr int XT = Cfg::XT, XN = Cfg::XN, NV = Cfg::NV, CPR = Cfg::CPR, NCP = Cfg::NCP;
    using Vec = VecT<T>;
    if (it > ctrl->converged_at) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *s_in = reinterpret_cast<T *>(smem_raw);  // [2][PH][PW]
    T *s_x = s_in + 2 * PH * PW;                // [PH][TX]

    const int tid = threadIdx.x;
    const int comp = blockIdx.z / nzchunks;
    const int chunk = blockIdx.z - comp * nzchunks;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z_begin = chunk * zchunk;
    const int z_end = min(z_begin + zchunk, g.S);
    const long long plane = static_cast<long long>(g.R) * g.C;
    const T *src = in + comp * g.N;
    T *dst = out + comp * g.N;

    // weights in registers; the host builds them symmetric bit for bit (w[k] == w[2R-k]), so R+1 per axis suffice
    T wxh[R + 1], wyh[R + 1], wzh[R + 1];
#pragma unroll
    for (int k = 0; k <= R; ++k) {
        wxh[k] = W.w[0][k];
        wyh[k] = W.w[1][k];
        wzh[k] = W.w[2][k];
    }
#define DCMA_WX(k) wxh[((k) <= R) ? (k) : (2 * R - (k))]
#define DCMA_WY(k) wyh[((k) <= R) ? (k) : (2 * R - (k))]
#define DCMA_WZ(k) wzh[((k) <= R) ? (k) : (2 * R - (k))]

    // ---- per-thread staging plan, fixed across slices: source offset inside a slice (-1: outside the grid -> zeros)
    int cp_src[NCP], cp_dst[NCP];
    if (VEC) {
#pragma unroll
        for (int n = 0; n < NCP; ++n) {
            const int e = tid + n * NT;
            const int row = e / CPR, cc = e - row * CPR;
            const int gy = y0 - R + row, gx = x0 - HX + cc * VW;
            const bool ok = (gy >= 0) && (gy < g.R) && (gx >= 0) && (gx + VW <= g.C);
            cp_dst[n] = (e < PH * CPR) ? (row * PW + cc * VW) : -1;
            cp_src[n] = ok ? (gy * g.C + gx) : -1;
        }
    }
    auto stage = [&](int zp, int buf) {
        if (zp < 0 || zp >= g.S) return;  // slice outside the grid: contributes zeros, nothing to load
        T *sb = s_in + buf * PH * PW;
        const T *sp = src + zp * plane;
        if (VEC) {
#pragma unroll
            for (int n = 0; n < NCP; ++n) {
                if (cp_dst[n] >= 0) cp_async_16(sb + cp_dst[n], (cp_src[n] >= 0) ? (sp + cp_src[n]) : src, (cp_src[n] >= 0) ? 16 : 0);
            }
        } else {
            for (int e = tid; e < PH * PW; e += NT) {
                const int row = e / PW, col = e - row * PW;
                const int gy = y0 - R + row, gx = x0 - HX + col;
                const bool ok = (gy >= 0) && (gy < g.R) && (gx >= 0) && (gx < g.C);
                cp_async_small<sizeof(T)>(sb + row * PW + col, ok ? (sp + static_cast<long long>(gy) * g.C + gx) : src,
                                          ok ? static_cast<int>(sizeof(T)) : 0);
            }
        }
    };

    // ---- thread geometry: column vector qx (x-pass, y-pass and outputs all use the same one), thread row ry
    const int qx = tid & (QX - 1), ry = tid / QX;
    const int gx_out = x0 + qx * VW;
    const int yl0 = ry * OY;  // first owned (local) output row; owned rows are adjacent
    T nx[VW], ny[OY];          // per-position normalisers (wsum or 1/wsum), fixed across slices
#pragma unroll
    for (int i = 0; i < VW; ++i) nx[i] = tab_x[gx_out + i];
#pragma unroll
    for (int j = 0; j < OY; ++j) ny[j] = tab_y[y0 + yl0 + j];

    // z window: the last TAPS xy-smoothed values of each owned output, in registers, used as a RING: the slice loop
    // is unrolled TAPS times so that every ring index is a compile-time constant (no register shuffling per slice).
    T win[OY][VW][TAPS];
#pragma unroll
    for (int j = 0; j < OY; ++j)
#pragma unroll
        for (int i = 0; i < VW; ++i)
#pragma unroll
            for (int k = 0; k < TAPS; ++k) win[j][i][k] = T(0);

    const int zp_first = z_begin - R, zp_last = z_end - 1 + R;
    stage(zp_first, 0);
    cp_async_commit();

    // one slice step; P = (zp - zp_first) % TAPS is the ring slot that receives this slice's xy-smoothed values
    auto step = [&](auto phase, const int zp) {
        constexpr int P = decltype(phase)::value;
        const int buf = (zp - zp_first) & 1;
        cp_async_wait_all();
        __syncthreads();  // s_in[buf] complete; everyone is done with s_x and s_in[buf^1] of the previous step
        if (zp + 1 <= zp_last) stage(zp + 1, buf ^ 1);
        cp_async_commit();

        const bool slice_in = (zp >= 0 && zp < g.S);
        if (slice_in) {
            const T *sb = s_in + buf * PH * PW;
            // ---- x pass: s_in -> s_x.  Task = one vector of VW outputs; q == qx for every task of this thread.
#pragma unroll
            for (int n = 0; n < XN; ++n) {
                const int row = ry + n * (NT / QX);
                if ((XT % NT != 0) && row >= PH) break;
                T a[VW * NV];
#pragma unroll
                for (int m = 0; m < NV; ++m) {
                    const Vec t = *reinterpret_cast<const Vec *>(sb + row * PW + (qx + m) * VW);
#pragma unroll
                    for (int i = 0; i < VW; ++i) a[m * VW + i] = t.v[i];
                }
                Vec o;
#pragma unroll
                for (int i = 0; i < VW; ++i) {
                    T sum = DCMA_WX(0) * a[HX - R + i];
#pragma unroll
                    for (int k = 1; k < TAPS; ++k) sum += DCMA_WX(k) * a[HX - R + i + k];
                    o.v[i] = normalise<T, STRICT>(sum, nx[i]);
                }
                *reinterpret_cast<Vec *>(s_x + row * TX + qx * VW) = o;
            }
            __syncthreads();
            // ---- y pass: s_x -> ring slot P.  Rows are streamed once; output j takes row m as its tap k = m - j
            // (ascending k: the reference's order); OY adjacent rows share their taps.
            T acc[OY][VW];
#pragma unroll
            for (int m = 0; m < OY + 2 * R; ++m) {
                const Vec t = *reinterpret_cast<const Vec *>(s_x + (yl0 + m) * TX + qx * VW);
#pragma unroll
                for (int j = 0; j < OY; ++j) {
                    const int k = m - j;
                    if (k < 0 || k >= TAPS) continue;
#pragma unroll
                    for (int i = 0; i < VW; ++i) {
                        if (k == 0)
                            acc[j][i] = DCMA_WY(0) * t.v[i];
                        else
                            acc[j][i] += DCMA_WY(k) * t.v[i];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < OY; ++j)
#pragma unroll
                for (int i = 0; i < VW; ++i) win[j][i][P] = normalise<T, STRICT>(acc[j][i], ny[j]);
        } else {
#pragma unroll
            for (int j = 0; j < OY; ++j)
#pragma unroll
                for (int i = 0; i < VW; ++i) win[j][i][P] = T(0);
        }
        // ---- z pass from the register ring: tap k (oldest first) lives in slot (P + 1 + k) % TAPS
        const int zo = zp - R;  // output slice completed by this step
        if (zo >= z_begin) {    // zo < z_end holds by construction of zp_last
            const T nz = tab_z[zo];
#pragma unroll
            for (int j = 0; j < OY; ++j) {
                const int gy = y0 + yl0 + j;
                Vec o;
                bool bad = false;
#pragma unroll
                for (int i = 0; i < VW; ++i) {
                    T sum = DCMA_WZ(0) * win[j][i][(P + 1) % TAPS];
#pragma unroll
                    for (int k = 1; k < TAPS; ++k) sum += DCMA_WZ(k) * win[j][i][(P + 1 + k) % TAPS];
                    o.v[i] = normalise<T, STRICT>(sum, nz);
                    bad = bad || !isfinite(o.v[i]);
                }
                if (bad) {
                    // a non-finite result <=> a non-finite tap somewhere in the (2r+1)^3 support: redo it literally
#pragma unroll
                    for (int i = 0; i < VW; ++i)
                        if (!isfinite(o.v[i]) && gy < g.R && gx_out + i < g.C)
                            o.v[i] = smooth_literal<T, R>(src, g, dwx, dwy, dwz, gx_out + i, gy, zo);
                }
                if (gy < g.R) {
                    T *p = dst + zo * plane + static_cast<long long>(gy) * g.C + gx_out;
                    if (VEC) {
                        if (gx_out < g.C) *reinterpret_cast<Vec *>(p) = o;  // C % VW == 0 in the VEC variant
                    } else {
#pragma unroll
                        for (int i = 0; i < VW; ++i)
                            if (gx_out + i < g.C) p[i] = o.v[i];
                    }
                }
            }
        }
    };

    for (int zb = zp_first; zb <= zp_last; zb += TAPS) unroll_phases<TAPS, 0>(step, zb, zp_last);
}

#undef DCMA_WX
#undef DCMA_WY
#undef DCMA_WZ

}  // namespace DCMA_NS
