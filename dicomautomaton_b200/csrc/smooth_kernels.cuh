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
//                       used for radii above kMaxFastRadius and as the `fused=0` path).
//  * k_smooth3d      -- ONE launch for all three passes: a CTA owns an x-y tile, marches along z, stages each input
//                       slice tile (+halo) in shared memory with cp.async, does the x and y passes in shared memory and
//                       keeps the last 2r+1 xy-smoothed values of its voxels in REGISTERS (a ring with compile-time
//                       indices) for the z pass.  Every field element is read from HBM once (+halo, served by L2) and
//                       written once.  Per-axis radii may differ (zero-padded weights, see SmoothWeights).
//                       fp32: packed FFMA2 lanes, one 128-thread role per CTA, 3 CTAs/SM.
//                       fp64: warp-specialised CTAs -- a producer warpgroup (staging + x pass) and a consumer warpgroup
//                       (y pass, ring, z pass, stores) with setmaxnreg and named barriers, 2 CTAs/SM.
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

// Tuning knobs of the fused kernel (tools/build_variant.py overrides them for experiments).
#ifndef DCMA_SMOOTH_NT
#define DCMA_SMOOTH_NT 128
#endif
#ifndef DCMA_SMOOTH_TY
#define DCMA_SMOOTH_TY 16
#endif
#ifndef DCMA_SMOOTH_MINB
#define DCMA_SMOOTH_MINB 3
#endif
#ifndef DCMA_SMOOTH_XPRELOAD
#define DCMA_SMOOTH_XPRELOAD 0
#endif
#ifndef DCMA_SMOOTH_WS
#define DCMA_SMOOTH_WS -1  // warp-specialised CTAs (producer/consumer warpgroups): 1 on, 0 off, -1 = per type (fp64 on:
                           // measured -8 % per launch at 512x512x256; fp32 off: no gain)
#endif
#ifndef DCMA_SMOOTH_XW
#define DCMA_SMOOTH_XW 1  // warp-specialised fp64 only: output vectors per x-pass task (1, 2 or 4).  Wider tasks read each
                          // staged value fewer times (shared-memory wavefronts of the x pass: 288 -> 192 -> 144 per step)
                          // but keep fewer producer threads busy: measured 4.11 / 4.29 / 4.62 ms per iteration for 1 / 2 / 4
#endif
#ifndef DCMA_SMOOTH_WS_PREG
#define DCMA_SMOOTH_WS_PREG 56   // registers per producer thread after setmaxnreg.dec
#endif
#ifndef DCMA_SMOOTH_WS_CREG
#define DCMA_SMOOTH_WS_CREG 200  // registers per consumer thread after setmaxnreg.inc (PREG + CREG <= 256)
#endif
#ifndef DCMA_SMOOTH_WS_JB
#define DCMA_SMOOTH_WS_JB 0  // warp-specialised fp64: advance all rows' z chains together (consumers have the registers)
#endif
#ifndef DCMA_SMOOTH_YBATCH
#define DCMA_SMOOTH_YBATCH 0
#endif
#ifndef DCMA_SMOOTH_PF
#define DCMA_SMOOTH_PF 0  // slices ahead of the cp.async stage that are requested into L2 (0 = off)
#endif

// Per-launch constants of one smoothing call (kernel parameter space => uniform/constant-bank reads).
// The fused kernel is instantiated for ONE radius R = max(r_x, r_y, r_z); an axis with a smaller radius gets its
// weights zero-padded around the centre (w[a][k], k - R = tap offset).  A zero-weight tap adds exactly 0 to the sum and
// to the weight sum, so the arithmetic of the true taps, order included, is unchanged.
template <typename T>
struct SmoothWeights {
    T w[3][2 * kMaxFastRadius + 1];  // [axis: 0=x,1=y,2=z][k + R]
    int r[3];                        // true radii per axis (cc:574-576), for the literal fallback
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
    const int extent = (axis == 0) ? g.C : (axis == 1) ? g.R : g.Sg;  // the z rule uses GLOBAL slice indices
    const long long stride = (axis == 0) ? 1 : (axis == 1) ? g.C : static_cast<long long>(g.R) * g.C;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < 3 * N; i += stride_thr) {
        const long long v = i % N;
        const int x = static_cast<int>(v % g.C);
        const long long r1 = v / g.C;
        const int y = static_cast<int>(r1 % g.R);
        const int z = static_cast<int>(r1 / g.R);
        // x and y passes run on every local slice (halo planes included: the z pass reads them); the z pass only on
        // the owned slices
        if (axis == 2 && (z < g.zb || z >= g.ze)) continue;
        const int pos = (axis == 0) ? x : (axis == 1) ? y : z + g.zoff;
        T sum = 0, wsum = 0;
        for (int k = -rad; k <= rad; ++k) {
            const int q = pos + k;
            if (q >= 0 && q < extent && (axis != 2 || (z + k >= 0 && z + k < g.S))) {
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

// ------------------------------------------------------------------------------------------------------------------
// Lane groups of the y and z passes.  fp32 (never STRICT): two x-adjacent outputs travel as ONE 64-bit register pair and
// every tap is one packed FFMA2 (PTX fma.rn.f32x2, sm_100+) with the weight broadcast to both lanes -- the same
// round-to-nearest FMA per lane as the scalar code, half the issue slots.  fp64 / STRICT: one output per group, plain
// `sum += w * v` so that the translation unit's -fmad setting decides about contraction.
// ------------------------------------------------------------------------------------------------------------------
template <typename T, bool PACKED>
struct LaneGroup {
    using V = T;
    static constexpr int L = 1;
    template <typename VecLike>
    static __device__ __forceinline__ V get(const VecLike &t, int p) { return t.v[p]; }
    template <typename VecLike>
    static __device__ __forceinline__ void put(VecLike &t, int p, V v) { t.v[p] = v; }
    static __device__ __forceinline__ V zero() { return T(0); }
    static __device__ __forceinline__ V pack(T lo, T) { return lo; }             // unused when L == 1
    static __device__ __forceinline__ void unpack(V v, T &lo, T &hi) { lo = hi = v; }  // unused when L == 1
    static __device__ __forceinline__ V mul2(V a, V b) { return a * b; }
    static __device__ __forceinline__ V mul(V a, T w) { return w * a; }
    static __device__ __forceinline__ V fma(V acc, V a, T w) {
        acc += w * a;
        return acc;
    }
};
template <>
struct LaneGroup<float, true> {
    using V = unsigned long long;  // {lo = element 2p, hi = element 2p+1}
    static constexpr int L = 2;
    static __device__ __forceinline__ V pack(float lo, float hi) {
        V r;
        asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
        return r;
    }
    template <typename VecLike>
    static __device__ __forceinline__ V get(const VecLike &t, int p) { return pack(t.v[2 * p], t.v[2 * p + 1]); }
    template <typename VecLike>
    static __device__ __forceinline__ void put(VecLike &t, int p, V v) {
        asm("mov.b64 {%0, %1}, %2;" : "=f"(t.v[2 * p]), "=f"(t.v[2 * p + 1]) : "l"(v));
    }
    static __device__ __forceinline__ V zero() { return 0ull; }
    static __device__ __forceinline__ void unpack(V v, float &lo, float &hi) {
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    }
    static __device__ __forceinline__ V mul2(V a, V b) {  // lane-wise product of two pairs
        V r;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
        return r;
    }
    static __device__ __forceinline__ V mul(V a, float w) {
        V r;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(pack(w, w)));
        return r;
    }
    static __device__ __forceinline__ V fma(V acc, V a, float w) {
        V r;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(pack(w, w)), "l"(acc));
        return r;
    }
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

// named barriers (ids 1..15; 0 is __syncthreads): `count` threads must arrive (bar.sync also waits)
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_pending() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

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
    // warp-specialised variant: 3 staged input tiles (two slice steps of lead) + 2 x-pass output tiles (hand-off)
    static constexpr size_t smem_bytes_ws = (size_t(3) * PH * PW + size_t(2) * PH * TX) * sizeof(T);
    static_assert(TX % VW == 0 && NT % QX == 0 && TY % RG == 0 && (QX & (QX - 1)) == 0, "tile/thread shape");
};

// The literal definition for ONE output voxel, straight from global memory (cc:609-693): used only when the fast path
// produced a non-finite value, i.e. when some tap in the (2r+1)^3 support is non-finite.  Deliberately not inlined.
template <typename T>
__device__ __noinline__ T smooth_literal(const T *__restrict__ src, const Geo g, const double *__restrict__ dwx,
                                         const double *__restrict__ dwy, const double *__restrict__ dwz, const int rx,
                                         const int ry, const int rz, int x, int y, int z) {
    const long long plane = static_cast<long long>(g.R) * g.C;
    T zs = 0, zw = 0, z_centre = 0;
    for (int kz = -rz; kz <= rz; ++kz) {
        const int zz = z + kz;  // local slice; in-grid test on the global index
        if (zz + g.zoff < 0 || zz + g.zoff >= g.Sg || zz < 0 || zz >= g.S) continue;
        T ys = 0, yw = 0, y_centre = 0;
        for (int ky = -ry; ky <= ry; ++ky) {
            const int yy = y + ky;
            if (yy < 0 || yy >= g.R) continue;
            const T *row = src + zz * plane + static_cast<long long>(yy) * g.C;
            T xs = 0, xw = 0;
            for (int kx = -rx; kx <= rx; ++kx) {
                const int xx = x + kx;
                if (xx < 0 || xx >= g.C) continue;
                const T val = row[xx];
                if (isfinite(val)) {
                    const T wk = static_cast<T>(dwx[kx + rx]);
                    xs += wk * val;
                    xw += wk;
                }
            }
            const T xv = (xw > T(0)) ? (xs / xw) : row[x];
            if (ky == 0) y_centre = xv;
            if (isfinite(xv)) {
                const T wk = static_cast<T>(dwy[ky + ry]);
                ys += wk * xv;
                yw += wk;
            }
        }
        const T yv = (yw > T(0)) ? (ys / yw) : y_centre;
        if (kz == 0) z_centre = yv;
        if (isfinite(yv)) {
            const T wk = static_cast<T>(dwz[kz + rz]);
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

// calls f(integral_constant<P>{}) for the run-time ring phase `phase` in [0, N): one uniform jump per slice, so that the
// slice step's common code (staging, x and y passes) exists once and only the ring insert + z pass are per phase.
template <int N, typename F>
__device__ __forceinline__ void dispatch_phase(const int phase, F &f) {
#define DCMA_PHASE_CASE(P)                                        \
    case P:                                                       \
        if constexpr (P < N) f(std::integral_constant<int, P>{}); \
        break;
    switch (phase) {
        DCMA_PHASE_CASE(0)
        DCMA_PHASE_CASE(1)
        DCMA_PHASE_CASE(2)
        DCMA_PHASE_CASE(3)
        DCMA_PHASE_CASE(4)
        DCMA_PHASE_CASE(5)
        DCMA_PHASE_CASE(6)
        DCMA_PHASE_CASE(7)
        DCMA_PHASE_CASE(8)
        DCMA_PHASE_CASE(9)
        DCMA_PHASE_CASE(10)
        DCMA_PHASE_CASE(11)
        DCMA_PHASE_CASE(12)
        default: break;
    }
#undef DCMA_PHASE_CASE
    static_assert(N <= 13, "ring too long for dispatch_phase");
}

// grid: x = tiles along columns, y = tiles along rows, z = 3 components * z-chunks.
// tab_x / tab_y / tab_z are padded by the host to whole tiles (no bounds tests here).
// WS (warp-specialised, DCMA_SMOOTH_WS): the CTA has 2*NT threads; threads [0, NT) are PRODUCERS (cp.async staging + x
// pass, ~56 registers after setmaxnreg.dec), threads [NT, 2*NT) CONSUMERS (y pass into the register ring, z pass, stores;
// ~200 registers after setmaxnreg.inc).  Hand-off through two x-pass tiles and named barriers: full[b] (producers
// arrive, consumers wait) and empty[b] (consumers arrive, producers wait).
template <typename T, bool STRICT, int R, int TX, int TY, int NT, int MINB, bool VEC, bool WS = false>
__global__ void __launch_bounds__(WS ? 2 * NT : NT, MINB)
    k_smooth3d(const T *__restrict__ in, T *__restrict__ out, const Geo g, const SmoothWeights<T> W,
               const T *__restrict__ tab_x, const T *__restrict__ tab_y, const T *__restrict__ tab_z,
               const double *__restrict__ dwx, const double *__restrict__ dwy, const double *__restrict__ dwz,
               const int zchunk, const int nzchunks, const Ctrl *__restrict__ ctrl, const long long it) {
    using Cfg = Smooth3dCfg<T, R, TX, TY, NT>;
    constexpr int VW = Cfg::VW, HX = Cfg::HX, PW = Cfg::PW, PH = Cfg::PH, QX = Cfg::QX, OY = Cfg::OY, TAPS = Cfg::TAPS;
    constexpr int XT = Cfg::XT, XN = Cfg::XN, NV = Cfg::NV, CPR = Cfg::CPR, NCP = Cfg::NCP;
    using Vec = VecT<T>;
    // fp32 (never STRICT): packed two-lane arithmetic and ONE deferred normalisation per pass pair (see below)
    constexpr bool PACKED = (!STRICT && sizeof(T) == 4);
    using LG = LaneGroup<T, PACKED>;
    using GV = typename LG::V;
    constexpr int NG = VW / LG::L;  // lane groups per 16-byte vector
    if (it > ctrl->converged_at) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *s_in = reinterpret_cast<T *>(smem_raw);  // [2 (WS: 3)][PH][PW]
    T *s_x = s_in + (WS ? 3 : 2) * PH * PW;     // [1 (WS: 2)][PH][TX]

    const bool producer = WS && (threadIdx.x < NT);
    const int tid = WS ? (threadIdx.x % NT) : threadIdx.x;
    const int comp = blockIdx.z / nzchunks;
    const int chunk = blockIdx.z - comp * nzchunks;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int z_begin = g.zb + chunk * zchunk;
    const int z_end = min(z_begin + zchunk, g.ze);
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
    // (16-byte chunk e = tid + n*NT of the staged tile lands at element offset e*VW, because PW == CPR*VW)
    int cp_src[NCP];
    if (VEC) {
#pragma unroll
        for (int n = 0; n < NCP; ++n) {
            const int e = tid + n * NT;
            const int row = e / CPR, cc = e - row * CPR;
            const int gy = y0 - R + row, gx = x0 - HX + cc * VW;
            const bool ok = (gy >= 0) && (gy < g.R) && (gx >= 0) && (gx + VW <= g.C);
            cp_src[n] = ok ? (gy * g.C + gx) : -1;
        }
    }
    static_assert(PW == CPR * VW, "staged row width must be a whole number of 16-byte chunks");
    auto l2_prefetch = [&](int zp) {  // one request per 128-byte line of the slice tile
        if (DCMA_SMOOTH_PF <= 0 || !VEC || zp < 0 || zp >= g.S) return;
        const T *sp = src + zp * plane;
#pragma unroll
        for (int n = 0; n < NCP; ++n) {
            if (cp_src[n] >= 0 && ((cp_src[n] * static_cast<int>(sizeof(T))) & 127) == 0)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + cp_src[n]));
        }
    };
    auto stage = [&](int zp, int buf) {
        // slice outside the GLOBAL grid: contributes zeros, nothing to load (a slab's halo planes are inside it)
        if (zp + g.zoff < 0 || zp + g.zoff >= g.Sg || zp < 0 || zp >= g.S) return;
        T *sb = s_in + buf * PH * PW;
        const T *sp = src + zp * plane;
        if (VEC) {
#pragma unroll
            for (int n = 0; n < NCP; ++n) {
                const int e = tid + n * NT;
                if ((n + 1) * NT <= PH * CPR || e < PH * CPR)
                    cp_async_16(sb + e * VW, (cp_src[n] >= 0) ? (sp + cp_src[n]) : src, (cp_src[n] >= 0) ? 16 : 0);
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
    // per-position normalisers (wsum, or 1/wsum when not STRICT), fixed across slices.  PACKED: the x and y
    // normalisers are applied together, once, after the y pass (they do not depend on the y or z tap): nxy = nx * ny.
    T nx[VW], ny[OY];
    GV nxy[OY][NG];
#pragma unroll
    for (int i = 0; i < VW; ++i) nx[i] = tab_x[gx_out + i];
#pragma unroll
    for (int j = 0; j < OY; ++j) ny[j] = tab_y[y0 + yl0 + j];
    if constexpr (PACKED) {
#pragma unroll
        for (int j = 0; j < OY; ++j)
#pragma unroll
            for (int i = 0; i < NG; ++i) nxy[j][i] = LG::pack(nx[2 * i] * ny[j], nx[2 * i + 1] * ny[j]);
    }
    // wide x-pass tasks (producers of the warp-specialised fp64 kernel): XW adjacent output vectors per task; a thread
    // keeps its column group, so the x normalisers of its XW*VW outputs are loop invariants
    constexpr int XW = (WS && !PACKED) ? DCMA_SMOOTH_XW : 1;
    static_assert(XW >= 1 && QX % XW == 0 && NT % (QX / XW) == 0, "x-pass task width");
    constexpr int XCG = QX / XW;   // column groups per staged row
    constexpr int XRPP = NT / XCG;  // rows covered per pass of the role's NT threads
    const int xcg = tid % XCG, xr0 = tid / XCG;
    T nxw[XW * VW];
    if constexpr (XW > 1) {
#pragma unroll
        for (int j = 0; j < XW * VW; ++j) nxw[j] = tab_x[x0 + xcg * XW * VW + j];
    }

    // z window: the last TAPS xy-smoothed values of each owned output, in registers, used as a RING: the ring insert
    // and the z pass exist once per ring phase so that every ring index is a compile-time constant.
    GV win[OY][NG][TAPS];
#pragma unroll
    for (int j = 0; j < OY; ++j)
#pragma unroll
        for (int i = 0; i < NG; ++i)
#pragma unroll
            for (int k = 0; k < TAPS; ++k) win[j][i][k] = LG::zero();

    const int zp_first = z_begin - R, zp_last = z_end - 1 + R;
    const auto in_grid = [&](int zp) { return zp + g.zoff >= 0 && zp + g.zoff < g.Sg && zp >= 0 && zp < g.S; };

    // ---- x pass of one staged slice tile: sb -> sx
    const T *sxr = s_x;  // the x-pass tile the y pass of the current step reads
    auto xpass = [&](const T *__restrict__ sb, T *__restrict__ sx) {
        {
            // ---- x pass: s_in -> s_x.  Task = one vector of VW outputs; q == qx for every task of this thread.
            // fp32: the operands of ALL x tasks are requested before the first task computes (the three shared-memory
            // latencies overlap instead of adding up; fp64 has no registers to spare for that)
            if constexpr (XW > 1) {
                constexpr int NVX = XW + 2 * HX / VW;  // vectors read per task
#pragma unroll
                for (int pass = 0; pass < (PH + XRPP - 1) / XRPP; ++pass) {
                    const int row = xr0 + pass * XRPP;
                    if (row >= PH) break;
                    T a[VW * NVX];
#pragma unroll
                    for (int m = 0; m < NVX; ++m) {
                        const Vec t = *reinterpret_cast<const Vec *>(sb + row * PW + (xcg * XW + m) * VW);
#pragma unroll
                        for (int i = 0; i < VW; ++i) a[m * VW + i] = t.v[i];
                    }
#pragma unroll
                    for (int v = 0; v < XW; ++v) {
                        Vec o;
#pragma unroll
                        for (int i = 0; i < VW; ++i) {
                            T sum = DCMA_WX(0) * a[HX - R + v * VW + i];  // the same taps in the same order as the narrow task
#pragma unroll
                            for (int k = 1; k < TAPS; ++k) sum += DCMA_WX(k) * a[HX - R + v * VW + i + k];
                            o.v[i] = normalise<T, STRICT>(sum, nxw[v * VW + i]);
                        }
                        *reinterpret_cast<Vec *>(sx + row * TX + (xcg * XW + v) * VW) = o;
                    }
                }
                return;
            }
            constexpr bool PRELOAD = PACKED && (DCMA_SMOOTH_XPRELOAD != 0);
            Vec pre[PRELOAD ? XN : 1][PRELOAD ? NV : 1];
            if constexpr (PRELOAD) {
#pragma unroll
                for (int n = 0; n < XN; ++n) {
                    const int row = ry + n * (NT / QX);
                    if ((XT % NT != 0) && row >= PH) break;
#pragma unroll
                    for (int m = 0; m < NV; ++m) pre[n][m] = *reinterpret_cast<const Vec *>(sb + row * PW + (qx + m) * VW);
                }
            }
#pragma unroll
            for (int n = 0; n < XN; ++n) {
                const int row = ry + n * (NT / QX);
                if ((XT % NT != 0) && row >= PH) break;
                T a[VW * NV];
#pragma unroll
                for (int m = 0; m < NV; ++m) {
                    Vec t;
                    if constexpr (PRELOAD)
                        t = pre[n][m];
                    else
                        t = *reinterpret_cast<const Vec *>(sb + row * PW + (qx + m) * VW);
#pragma unroll
                    for (int i = 0; i < VW; ++i) a[m * VW + i] = t.v[i];
                }
                Vec o;
                if constexpr (PACKED) {
                    // Packed x pass.  Aligned register pairs A[m] = (a[2m], a[2m+1]).  Output pair p = (out[2p],
                    // out[2p+1]) takes tap k from (a[off+2p+k], a[off+2p+k+1]), an aligned pair only when off+k is
                    // even: those taps go to E[p].  The others are accumulated for the SHIFTED output pairs
                    // O[q] = (out[2q-1], out[2q]), whose operands (a[off+2q-1+k], a[off+2q+k]) are aligned, and
                    // out[2p] = E[p].lo + O[p].hi, out[2p+1] = E[p].hi + O[p+1].lo.  Unnormalised (see nxy).
                    constexpr int off = HX - R, NP = VW * NV / 2;
                    GV A[NP];
#pragma unroll
                    for (int m = 0; m < NP; ++m) A[m] = LG::pack(a[2 * m], a[2 * m + 1]);
                    GV E[VW / 2], O[VW / 2 + 1];
#pragma unroll
                    for (int k = 0; k < TAPS; ++k) {
                        if (((off + k) & 1) == 0) {
#pragma unroll
                            for (int pp = 0; pp < VW / 2; ++pp) {
                                const int m = (off + k) / 2 + pp;
                                E[pp] = (k == ((off & 1) ? 1 : 0)) ? LG::mul(A[m], DCMA_WX(k)) : LG::fma(E[pp], A[m], DCMA_WX(k));
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q <= VW / 2; ++q) {
                                const int m = (off + k - 1) / 2 + q;  // pair (a[off+k-1+2q], a[off+k+2q])
                                const GV Am = (m >= 0 && m < NP) ? A[m] : LG::zero();
                                O[q] = (k == ((off & 1) ? 0 : 1)) ? LG::mul(Am, DCMA_WX(k)) : LG::fma(O[q], Am, DCMA_WX(k));
                            }
                        }
                    }
#pragma unroll
                    for (int pp = 0; pp < VW / 2; ++pp) {
                        float e_lo, e_hi, o_lo, o_hi, o2_lo, o2_hi;
                        LG::unpack(E[pp], e_lo, e_hi);
                        LG::unpack(O[pp], o_lo, o_hi);
                        LG::unpack(O[pp + 1], o2_lo, o2_hi);
                        o.v[2 * pp] = e_lo + o_hi;
                        o.v[2 * pp + 1] = e_hi + o2_lo;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < VW; ++i) {
                        T sum = DCMA_WX(0) * a[HX - R + i];
#pragma unroll
                        for (int k = 1; k < TAPS; ++k) sum += DCMA_WX(k) * a[HX - R + i + k];
                        o.v[i] = normalise<T, STRICT>(sum, nx[i]);
                    }
                }
                *reinterpret_cast<Vec *>(sx + row * TX + qx * VW) = o;
            }
        }
    };

    // One slice step (not WS) = xstep(zp) [staging + x pass, phase independent] followed by ring(P) [y pass into ring
    // slot P = (zp - zp_first) % TAPS, then the z pass].
    T nz_next = tab_z[z_begin < z_end ? z_begin : 0];
    bool slice_in = false;
    int zo = 0;               // output slice completed by the current step
    int ws_b2 = 0;            // WS: index of the x-pass tile of the current step
    bool ws_release = false;  // WS: whether the producers will wait for this tile's release
    auto xstep = [&](const int zp) {
        const int buf = (zp - zp_first) & 1;
        cp_async_wait_all();
        __syncthreads();  // s_in[buf] complete; everyone is done with s_x and s_in[buf^1] of the previous step
        if (zp + 1 <= zp_last) stage(zp + 1, buf ^ 1);
        cp_async_commit();
        if (zp + 1 + DCMA_SMOOTH_PF <= zp_last) l2_prefetch(zp + 1 + DCMA_SMOOTH_PF);
        slice_in = in_grid(zp);
        zo = zp - R;
        if (slice_in) {
            xpass(s_in + buf * PH * PW, s_x);
            __syncthreads();
        }
    };

    // ---- per ring phase P (compile-time ring indices): y pass straight into ring slot P, then the z pass.
    // Tap k of the z pass (oldest first) lives in slot (P + 1 + k) % TAPS.
    auto ring = [&](auto ph) {
        {
            constexpr int P = decltype(ph)::value;
            if (slice_in) {
                // y pass: s_x -> win[..][P].  Rows are streamed once; output j takes row m as its tap k = m - j
                // (ascending k: the reference's order); OY adjacent rows share their taps.
#pragma unroll
                for (int m = 0; m < OY + 2 * R; ++m) {
                    // DCMA_SMOOTH_YBATCH > 0: at most that many row vectors are in flight (a compiler-level fence
                    // keeps the shared-memory loads from all being hoisted to the front) -- fewer live registers
                    if (DCMA_SMOOTH_YBATCH > 0 && m > 0 && (m % DCMA_SMOOTH_YBATCH) == 0) asm volatile("" ::: "memory");
                    const Vec t = *reinterpret_cast<const Vec *>(sxr + (yl0 + m) * TX + qx * VW);
#pragma unroll
                    for (int j = 0; j < OY; ++j) {
                        const int k = m - j;
                        if (k < 0 || k >= TAPS) continue;
#pragma unroll
                        for (int i = 0; i < NG; ++i) {
                            if (k == 0)
                                win[j][i][P] = LG::mul(LG::get(t, i), DCMA_WY(0));
                            else
                                win[j][i][P] = LG::fma(win[j][i][P], LG::get(t, i), DCMA_WY(k));
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < OY; ++j)
#pragma unroll
                    for (int i = 0; i < NG; ++i) {
                        if constexpr (STRICT)
                            win[j][i][P] = win[j][i][P] / ny[j];  // STRICT => LG::V == T
                        else if constexpr (PACKED)
                            win[j][i][P] = LG::mul2(win[j][i][P], nxy[j][i]);
                        else
                            win[j][i][P] = LG::mul(win[j][i][P], ny[j]);
                    }
            } else {
#pragma unroll
                for (int j = 0; j < OY; ++j)
#pragma unroll
                    for (int i = 0; i < NG; ++i) win[j][i][P] = LG::zero();
            }
            if constexpr (WS) {  // the y pass has the tile in registers: the producers may overwrite it
                if (ws_release) named_bar_arrive(4 + ws_b2, 2 * NT);
            }
            // z normaliser of this output slice, requested one step ago (its L2 latency would otherwise sit on the
            // critical path of every step); the next one is requested now
            // (not in the plain fp64 kernel: it sits at its 168-register cap and the extra value would spill)
            constexpr bool NZ_AHEAD = WS || PACKED;
            T nz = nz_next;
            if constexpr (NZ_AHEAD) nz_next = tab_z[min(max(zo + 1, z_begin), z_end - 1)];
            if (zo < z_begin) return;  // warm-up slices of this chunk; zo < z_end holds by construction of zp_last
            if constexpr (!NZ_AHEAD) nz = tab_z[zo];
            // JB rows advance together, tap-major: independent FMA chains back to back (fp32: all rows; fp64 keeps
            // one row at a time, its register file has no room for more accumulators)
            constexpr int JB = (PACKED || (WS && DCMA_SMOOTH_WS_JB)) ? OY : 1;
#pragma unroll
            for (int jb = 0; jb < OY; jb += JB) {
                GV zs[JB][NG];
#pragma unroll
                for (int k = 0; k < TAPS; ++k)
#pragma unroll
                    for (int jj = 0; jj < JB; ++jj)
#pragma unroll
                        for (int i = 0; i < NG; ++i)
                            zs[jj][i] = (k == 0) ? LG::mul(win[jb + jj][i][(P + 1) % TAPS], DCMA_WZ(0))
                                                 : LG::fma(zs[jj][i], win[jb + jj][i][(P + 1 + k) % TAPS], DCMA_WZ(k));
#pragma unroll
                for (int jj = 0; jj < JB; ++jj) {
                const int j = jb + jj;
                const int gy = y0 + yl0 + j;
                Vec o;
                bool bad = false;
#pragma unroll
                for (int i = 0; i < NG; ++i) {
                    if constexpr (STRICT)
                        LG::put(o, i, zs[jj][i] / nz);
                    else
                        LG::put(o, i, LG::mul(zs[jj][i], nz));
                }
                // one finite test per vector: x*0 is 0 for finite x and NaN otherwise
                T chk = o.v[0] * T(0);
#pragma unroll
                for (int i = 1; i < VW; ++i) chk += o.v[i] * T(0);
                bad = !(chk == T(0));
                if (bad) {
                    // a non-finite result <=> a non-finite tap somewhere in the (2r+1)^3 support: redo it literally
#pragma unroll
                    for (int i = 0; i < VW; ++i)
                        if (!isfinite(o.v[i]) && gy < g.R && gx_out + i < g.C)
                            o.v[i] = smooth_literal<T>(src, g, dwx, dwy, dwz, W.r[0], W.r[1], W.r[2], gx_out + i, gy, zo);
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
        }
    };

    if constexpr (WS) {
        constexpr int kBarProd = 1, kBarFull = 2, kBarEmpty = 4;  // named barrier ids (+ tile index for full/empty)
        const int n_steps = zp_last - zp_first + 1;
        if (producer) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DCMA_SMOOTH_WS_PREG));
            stage(zp_first, 0);
            cp_async_commit();
            if (zp_first + 1 <= zp_last) stage(zp_first + 1, 1);
            cp_async_commit();
            int b3 = 0;
            for (int it2 = 0; it2 < n_steps; ++it2) {
                const int zp = zp_first + it2, b2 = it2 & 1;
                cp_async_wait_pending<1>();    // this thread's chunks of slice zp have landed (one newer group may fly)
                named_bar_sync(kBarProd, NT);  // ... and every producer's; all are done with the x pass of slice zp-1
                const int n3 = (b3 == 0) ? 2 : b3 - 1;  // (it2 + 2) % 3: the tile the x pass of slice zp-1 has just left
                if (zp + 2 <= zp_last) stage(zp + 2, n3);
                cp_async_commit();
                if (it2 >= 2) named_bar_sync(kBarEmpty + b2, 2 * NT);  // consumers are done with the tile of slice zp-2
                if (in_grid(zp)) xpass(s_in + b3 * PH * PW, s_x + b2 * PH * TX);
                named_bar_arrive(kBarFull + b2, 2 * NT);
                b3 = (b3 == 2) ? 0 : b3 + 1;
            }
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(DCMA_SMOOTH_WS_CREG));
        static_assert(DCMA_SMOOTH_WS_PREG + DCMA_SMOOTH_WS_CREG <= 256, "register split exceeds the CTA's allocation");
        if constexpr (PACKED) {
            int phase = 0;
            for (int it2 = 0; it2 < n_steps; ++it2) {
                const int zp = zp_first + it2;
                ws_b2 = it2 & 1;
                ws_release = (it2 + 2 < n_steps);  // the last two tiles are never reused: nobody waits for their release
                named_bar_sync(kBarFull + ws_b2, 2 * NT);
                sxr = s_x + ws_b2 * PH * TX;
                slice_in = in_grid(zp);
                zo = zp - R;
                dispatch_phase<TAPS>(phase, ring);
                phase = (phase + 1 == TAPS) ? 0 : phase + 1;
            }
        } else {
            auto step = [&](auto ph, const int zp) {
                const int it2 = zp - zp_first;
                ws_b2 = it2 & 1;
                ws_release = (it2 + 2 < n_steps);
                named_bar_sync(kBarFull + ws_b2, 2 * NT);
                sxr = s_x + ws_b2 * PH * TX;
                slice_in = in_grid(zp);
                zo = zp - R;
                ring(ph);
            };
            for (int zb = zp_first; zb <= zp_last; zb += TAPS) unroll_phases<TAPS, 0>(step, zb, zp_last);
        }
        return;
    }
    stage(zp_first, 0);
    cp_async_commit();
    if constexpr (PACKED) {
        // fp32: one copy of the step; a uniform jump per slice selects the ring phase (small code, fits the I-cache)
        int phase = 0;
        for (int zp = zp_first; zp <= zp_last; ++zp) {
            xstep(zp);
            dispatch_phase<TAPS>(phase, ring);
            phase = (phase + 1 == TAPS) ? 0 : phase + 1;
        }
    } else {
        // fp64: the whole step is unrolled per ring phase (measured 10-25 % faster than the dispatch form here: the
        // scheduler interleaves the x pass of a step with the ring arithmetic around it)
        auto step = [&](auto ph, const int zp) {
            xstep(zp);
            ring(ph);
        };
        for (int zb = zp_first; zb <= zp_last; zb += TAPS) unroll_phases<TAPS, 0>(step, zb, zp_last);
    }
}

#undef DCMA_WX
#undef DCMA_WY
#undef DCMA_WZ

}  // namespace DCMA_NS
