// dcma_b200 -- k_smooth3d_tma: the fused separable 3-D Gaussian of smooth_on_device (reference
// src/Alignment_Demons.cc:563-702) for fp64 fields, re-cut around what ncu showed to be the limit of k_smooth3d<double>:
// the shared-memory data pipe (LSU wavefronts 76 % of peak at ~1.0 wavefront per output), not HBM (37 %).
//
//   k_smooth3d<double> (smooth_kernels.cuh)                 this kernel
//   --------------------------------------------------      -----------------------------------------------------------
//   cp.async staging, 128 producer threads each with a      ONE elected thread issues one TMA box per slice
//   per-thread address plan (60 wavefronts per step)         (cp.async.bulk.tensor.4d; out-of-grid rows/columns are
//                                                            zero-filled by the TMA unit = the reference's "tap outside
//                                                            the grid contributes nothing"); full/empty mbarriers
//   x pass: one 16-byte vector of outputs per task,          x pass: a task = 8 adjacent outputs of one row, 16 inputs
//   5 vectors read per vector written (288 wavefronts)       read once (2.0 values read per output instead of 5.0); the
//                                                            32 lanes of a warp take 32 ROWS, row pitch = odd number of
//                                                            16-byte units => conflict-free LDS.128 / STS.128
//   y pass: 2 x 2 outputs per consumer, 10 rows streamed     y pass: 1 column x OY rows per consumer, OY + 2R rows
//   for 2 (160 wavefronts)                                   streamed for OY, lanes along x (conflict-free LDS.64)
//   z pass: register ring, compile-time phases               unchanged
//
// Shared-memory wavefronts per output: ~1.0 -> ~0.55-0.6 (R = 4, 32 x 24 tile).  Arithmetic per output is unchanged and in
// the SAME ORDER (x taps ascending, y taps ascending, z taps ascending; STRICT divides by the in-grid weight sum after each
// pass exactly as the reference does), so STRICT results stay bit-identical to the reference engine.  The non-STRICT mode
// applies the three per-position normalisers once, at the end (they factor out of the later passes), as the fp32 kernel does.
//
// CTA = 4 producer warps (x pass; warp 0 lane 0 also issues the TMA boxes) + TY/OY consumer warps (y pass into the register
// ring, z pass, stores).  setmaxnreg moves registers from the producers to the consumers.  2 CTAs per SM.
#pragma once
#include <cuda.h>  // CUtensorMap (type only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint)

#include "smooth_kernels.cuh"

#ifndef DCMA_TMA_TX
#define DCMA_TMA_TX 64  // tile columns: TX/8 producer warps (whole warpgroups: 32 or 64)
#endif
#ifndef DCMA_TMA_TY
#define DCMA_TMA_TY 24  // tile rows; TY + 2R = 32 staged rows for R = 4: one x-pass task row per lane
#endif
#ifndef DCMA_TMA_OY
#define DCMA_TMA_OY 4   // rows per consumer thread: (TX/32) * (TY/OY) consumer warps
#endif
#ifndef DCMA_TMA_NS
#define DCMA_TMA_NS 3   // TMA stages (slices in flight)
#endif
#ifndef DCMA_TMA_NXB
#define DCMA_TMA_NXB 2  // x-pass output tiles (producer -> consumer hand-off)
#endif
#ifndef DCMA_TMA_PREG
#define DCMA_TMA_PREG 48  // registers per producer thread after setmaxnreg.dec
#endif
#ifndef DCMA_TMA_CREG
#define DCMA_TMA_CREG 128  // registers per consumer thread after setmaxnreg.inc; see the pool static_assert below
#endif
#ifndef DCMA_TMA_MINB
#define DCMA_TMA_MINB (DCMA_TMA_TX == 32 ? 2 : 1)  // CTAs per SM
#endif
#ifndef DCMA_TMA_IDLE_WARPS
#define DCMA_TMA_IDLE_WARPS 0  // warps that only shrink to 24 registers and exit (pads the CTA to whole warpgroups)
#endif

namespace DCMA_NS {
using namespace ::dcma;

// ---- mbarrier / TMA primitives (PTX ISA 8.x; sm_90+; here sm_100a) -----------------------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// Spin on the phase parity.  A protocol error would otherwise hang the GPU until the driver's watchdog fires, so the spin
// is bounded (~seconds) and traps: the launch then fails with an error instead of wedging the device.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned addr = smem_addr(bar);
    unsigned done = 0;
    for (unsigned spins = 0; !done; ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && spins > (1u << 22)) __trap();
    }
}
// one box {c0 = x, c1 = y, c2 = z, c3 = component} of the 4-D tensor map -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, int c3,
                                            unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_addr(smem_dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(smem_addr(bar))
        : "memory");
}

// Outputs (x, y0..y0+ny-1, z0..z1-1) of one thread that the fast path left non-finite are recomputed with the literal
// per-tap definition (cc:609-693): taps that are non-finite are skipped and the weights renormalised.  Rare, cold, and
// deliberately one out-of-line call per thread so that the marching loop carries no call sites.
template <typename T>
__device__ __noinline__ void smooth_fix_column(const T *__restrict__ src, T *__restrict__ dst, const Geo g,
                                               const double *__restrict__ dwx, const double *__restrict__ dwy,
                                               const double *__restrict__ dwz, const int rx, const int ry, const int rz,
                                               const int x, const int y0, const int ny, const int z0, const int z1) {
    const long long plane = static_cast<long long>(g.R) * g.C;
    for (int z = z0; z < z1; ++z)
        for (int j = 0; j < ny; ++j) {
            const int y = y0 + j;
            if (y >= g.R) break;
            T *p = dst + z * plane + static_cast<long long>(y) * g.C + x;
            if (!isfinite(*p)) *p = smooth_literal<T>(src, g, dwx, dwy, dwz, rx, ry, rz, x, y, z);
        }
}

template <typename T, int R, int TX, int TY, int OY, int NS, int NXB>
struct SmoothTmaCfg {
    static_assert(sizeof(T) == 8, "the TMA smoother is the fp64 kernel");
    static constexpr int SEG = 8;                        // outputs per x-pass task
    static constexpr int NSEG = TX / SEG;                // = producer warps (one segment column per warp)
    static constexpr int HX = ((R + 1) / 2) * 2;         // x halo rounded up to 16 bytes
    static constexpr int NA = SEG + 2 * HX;              // values read per x-pass task
    static constexpr int PH = TY + 2 * R;                // staged rows
    static constexpr int PWP = TX + 2 * HX + 2;          // staged row pitch: odd number of 16-byte units
    static constexpr int SXP = TX + 2;                   // x-pass tile row pitch, same rule
    static constexpr int TAPS = 2 * R + 1;
    static constexpr int NPW = NSEG;                     // producer warps
    static constexpr int NCX = TX / 32;                  // consumer warps side by side (32 columns each)
    static constexpr int NCW = NCX * (TY / OY);          // consumer warps
    static constexpr int NIW = DCMA_TMA_IDLE_WARPS;      // idle warps (behind the consumers)
    static constexpr int NTHREADS = 32 * (NPW + NCW + NIW);
    static constexpr unsigned stage_bytes = PH * PWP * sizeof(T);                    // what one TMA box delivers
    static constexpr unsigned stage_stride = (stage_bytes + 127u) & ~127u;           // TMA destinations: 128-byte aligned
    static constexpr unsigned sx_stride = (PH * SXP * unsigned(sizeof(T)) + 127u) & ~127u;
    static constexpr unsigned off_sx = NS * stage_stride;
    static constexpr unsigned off_nx = off_sx + NXB * sx_stride;
    static constexpr unsigned off_bar = off_nx + TX * unsigned(sizeof(T));
    static constexpr size_t smem_bytes = off_bar + (2 * NS + 2 * NXB) * sizeof(unsigned long long);
    static_assert(TX % 32 == 0, "consumer lanes run along 32 columns of the tile");
    // setmaxnreg.sync.aligned is a WARPGROUP collective: every warpgroup (4 consecutive warps) must be complete and
    // execute the same instruction -- a partial group, or one split between roles, hangs (measured on B200)
    static_assert(NPW % 4 == 0 && (NCW + NIW) % 4 == 0 && (NIW == 0 || DCMA_TMA_CREG == 24), "roles must be whole warpgroups");
    static_assert(TY % OY == 0 && PH <= 32, "tile/thread shape");
    static_assert(((PWP / 2) & 1) == 1 && ((SXP / 2) & 1) == 1, "row pitches must be odd multiples of 16 bytes");
    static_assert(PWP <= 256 && PH <= 256, "TMA box limits");
    // setmaxnreg moves registers inside the pool the CTA was LAUNCHED with (threads x registers at launch, 8-register
    // granules); setmaxnreg.inc blocks until the pool can serve it, so an over-subscribed split hangs the kernel.
    static constexpr int launch_regs = (65536 / (NTHREADS * DCMA_TMA_MINB)) / 8 * 8;
    static_assert(32 * (NPW * DCMA_TMA_PREG + NCW * DCMA_TMA_CREG + NIW * 24) <= NTHREADS * launch_regs,
                  "register split exceeds the CTA's launch allocation: setmaxnreg.inc would never return");
};

// ---- work decomposition -------------------------------------------------------------------------------------------
// A "column" is one x-y tile of one component, marched along z.  The launch is PERSISTENT: G = (CTAs per SM) x (SMs)
// CTAs, and the (columns x owned slices) work is cut so that every CTA gets the same number of slice steps:
//   rounds 0 .. nfull-1 : CTA b marches the whole column r*G + b (adjacent CTAs -> adjacent tiles: halos are shared in L2);
//   the remaining columns: their slices, laid end to end, are cut into G equal pieces (a piece may span two columns).
// With equal tiles a plain grid loses up to a whole tile-time in its last wave (528 columns on 148 SMs: 3.57 -> 4 rounds);
// here the only overhead is the 2R warm-up slices at the head of each piece of the last round.
struct SmoothSeg {
    int col;     // column index = (comp * tiles_y + ty) * tiles_x + tx
    int z0, z1;  // owned output slices [z0, z1) (local indices)
};
struct SmoothWork {
    int ncols, tiles_x, tiles_y;  // columns = 3 * tiles_y * tiles_x
    int nfull;                    // whole-column rounds
    int per;                      // slices per CTA in the last round
    __device__ __forceinline__ int count(int bid, int G, int So) const {
        (void)G;
        const long long rem = static_cast<long long>(ncols - nfull * G) * So;
        const long long l0 = static_cast<long long>(bid) * per, l1 = min(l0 + per, rem);
        int n = nfull;
        if (l0 < l1) n += static_cast<int>((l1 - 1) / So - l0 / So) + 1;
        return n;
    }
    __device__ __forceinline__ SmoothSeg get(int k, int bid, int G, int zb, int So) const {
        SmoothSeg sg;
        if (k < nfull) {
            sg.col = k * G + bid;
            sg.z0 = zb;
            sg.z1 = zb + So;
            return sg;
        }
        const long long rem = static_cast<long long>(ncols - nfull * G) * So;
        const long long l0 = static_cast<long long>(bid) * per, l1 = min(l0 + per, rem);
        const int c0 = static_cast<int>(l0 / So) + (k - nfull);  // k-th column touched by [l0, l1)
        const long long a = max(l0, static_cast<long long>(c0) * So), b = min(l1, static_cast<long long>(c0 + 1) * So);
        sg.col = nfull * G + c0;
        sg.z0 = zb + static_cast<int>(a - static_cast<long long>(c0) * So);
        sg.z1 = zb + static_cast<int>(b - static_cast<long long>(c0) * So);
        return sg;
    }
};

template <typename T, bool STRICT, int R, int TX, int TY, int OY, int NS, int NXB>
__global__ void __launch_bounds__((SmoothTmaCfg<T, R, TX, TY, OY, NS, NXB>::NTHREADS), DCMA_TMA_MINB)
    k_smooth3d_tma(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ in, T *__restrict__ out, const Geo g,
                   const SmoothWeights<T> W, const T *__restrict__ tab_x, const T *__restrict__ tab_y,
                   const T *__restrict__ tab_z, const double *__restrict__ dwx, const double *__restrict__ dwy,
                   const double *__restrict__ dwz, const SmoothWork work, const Ctrl *__restrict__ ctrl,
                   const long long it) {
    using Cfg = SmoothTmaCfg<T, R, TX, TY, OY, NS, NXB>;
    constexpr int HX = Cfg::HX, NA = Cfg::NA, PH = Cfg::PH, PWP = Cfg::PWP, SXP = Cfg::SXP, TAPS = Cfg::TAPS, SEG = Cfg::SEG;
    constexpr int NPW = Cfg::NPW, NCW = Cfg::NCW;
    if (it > ctrl->converged_at) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *s_in = reinterpret_cast<T *>(smem_raw);                  // [NS][PH][PWP]  (stage_stride bytes apart)
    T *s_x = reinterpret_cast<T *>(smem_raw + Cfg::off_sx);     // [NXB][PH][SXP]
    T *s_nx = reinterpret_cast<T *>(smem_raw + Cfg::off_nx);    // [TX] x normalisers (STRICT: divided by after the x pass)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw + Cfg::off_bar);
    unsigned long long *full_in = bars, *empty_in = bars + NS, *full_x = bars + 2 * NS, *empty_x = bars + 2 * NS + NXB;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bid = blockIdx.x, G = gridDim.x;
    const int So = g.ze - g.zb;
    const long long plane = static_cast<long long>(g.R) * g.C;
    const int nseg = work.count(bid, G, So);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(full_in + s, 1);     // the arming arrive of the TMA thread (+ the box's bytes)
            mbar_init(empty_in + s, NPW);  // one arrive per producer warp
        }
#pragma unroll
        for (int b = 0; b < NXB; ++b) {
            mbar_init(full_x + b, NPW);
            mbar_init(empty_x + b, NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

#define DCMA_WX(k) W.w[0][k]
#define DCMA_WY(k) W.w[1][k]
#define DCMA_WZ(k) W.w[2][k]
    // Per segment: slices staged = [zp_first, zp_last] = [z0 - R, z1 - 1 + R]; those inside the grid AND inside this
    // engine's buffers form one contiguous range [zlo, zhi] (a slab's halo planes are inside it); the others contribute
    // zeros and are skipped by every role alike, so the pipeline counters count in-grid slices only -- and they run on
    // across segments (the barriers' phases do not know about segments).

    static_assert(DCMA_TMA_IDLE_WARPS == 0, "idle warps cannot share a warpgroup with consumers (setmaxnreg is warpgroup-wide)");
    if (warp < NPW) {
        // =============================================== producers =======================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DCMA_TMA_PREG));
        const bool tma_thread = (threadIdx.x == 0);
        const int seg = warp, row = lane;  // task: outputs [seg*8, seg*8+8) of staged row `row`
        int si = 0, phi = 1;                // TMA thread: next stage to fill and the parity of its `empty` wait
        int s = 0, ph_in = 0;               // x pass: stage being read and its `full` parity
        int b = 0, ph_x = 1;                // x-pass tile being written and the parity of its `empty` wait
        for (int k = 0; k < nseg; ++k) {
            const SmoothSeg sg = work.get(k, bid, G, g.zb, So);
            const int tx = sg.col % work.tiles_x, rest = sg.col / work.tiles_x;
            const int ty = rest % work.tiles_y, comp = rest / work.tiles_y;
            const int x0 = tx * TX, y0 = ty * TY;
            const int zp_first = sg.z0 - R, zp_last = sg.z1 - 1 + R;
            const int zlo = max(zp_first, max(0, -g.zoff));
            const int zhi = min(zp_last, min(g.S - 1, g.Sg - 1 - g.zoff));
            const int n_in = zhi - zlo + 1;
            if constexpr (STRICT) {  // this warp's 8 x normalisers (read back by this warp only)
                if (lane < SEG) s_nx[seg * SEG + lane] = tab_x[x0 + seg * SEG + lane];
                __syncwarp();
            }
            auto issue = [&](int z) {
                mbar_wait(empty_in + si, phi);  // first use of a stage: passes at once (fresh barrier, opposite parity)
                mbar_arrive_expect_tx(full_in + si, Cfg::stage_bytes);
                tma_load_4d(reinterpret_cast<unsigned char *>(s_in) + si * Cfg::stage_stride, &tmap, x0 - HX, y0 - R, z, comp,
                            full_in + si);
                if (++si == NS) {
                    si = 0;
                    phi ^= 1;
                }
            };
            if (tma_thread) {
                for (int i = 0; i < NS - 1 && i < n_in; ++i) issue(zlo + i);
            }
            for (int i = 0; i < n_in; ++i) {
                if (tma_thread && i + NS - 1 < n_in) issue(zlo + i + NS - 1);
                __syncwarp();
                mbar_wait(full_in + s, ph_in);
                mbar_wait(empty_x + b, ph_x);
                if (PH == 32 || row < PH) {
                    const T *sb = reinterpret_cast<const T *>(reinterpret_cast<const unsigned char *>(s_in) + s * Cfg::stage_stride) +
                                  row * PWP + seg * SEG;
                    T a[NA];
#pragma unroll
                    for (int m = 0; m < NA / 2; ++m) {
                        const double2 t = *reinterpret_cast<const double2 *>(sb + 2 * m);
                        a[2 * m] = t.x;
                        a[2 * m + 1] = t.y;
                    }
                    T o[SEG];
#pragma unroll
                    for (int i2 = 0; i2 < SEG; ++i2) {
                        T sum = DCMA_WX(0) * a[HX - R + i2];  // taps in ascending order, as the reference (cc:619-631)
#pragma unroll
                        for (int kk = 1; kk < TAPS; ++kk) sum += DCMA_WX(kk) * a[HX - R + i2 + kk];
                        o[i2] = sum;
                    }
                    if constexpr (STRICT) {
#pragma unroll
                        for (int m = 0; m < SEG / 2; ++m) {
                            const double2 n2 = *reinterpret_cast<const double2 *>(s_nx + seg * SEG + 2 * m);
                            o[2 * m] = o[2 * m] / n2.x;
                            o[2 * m + 1] = o[2 * m + 1] / n2.y;
                        }
                    }
                    T *sx = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(s_x) + b * Cfg::sx_stride) + row * SXP + seg * SEG;
#pragma unroll
                    for (int m = 0; m < SEG / 2; ++m) *reinterpret_cast<double2 *>(sx + 2 * m) = make_double2(o[2 * m], o[2 * m + 1]);
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(full_x + b);    // this warp's segment column of the x-pass tile is written
                    mbar_arrive(empty_in + s);  // ... and it is done reading the staged slice
                }
                if (++s == NS) {
                    s = 0;
                    ph_in ^= 1;
                }
                if (++b == NXB) {
                    b = 0;
                    ph_x ^= 1;
                }
            }
        }
        return;
    }

    // ================================================= consumers =========================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(DCMA_TMA_CREG));
    const int cw = warp - NPW;
    const int xh = cw % Cfg::NCX;          // which 32-column part of the tile
    const int yl0 = (cw / Cfg::NCX) * OY;  // first owned (local) output row; owned rows are adjacent
    int b = 0, ph_x = 0;                   // x-pass tile being read and its `full` parity
    for (int k = 0; k < nseg; ++k) {
        const SmoothSeg sg = work.get(k, bid, G, g.zb, So);
        const int tx = sg.col % work.tiles_x, rest = sg.col / work.tiles_x;
        const int ty = rest % work.tiles_y, comp = rest / work.tiles_y;
        const int x0 = tx * TX, y0 = ty * TY;
        const int z_begin = sg.z0, z_end = sg.z1;
        const int zp_first = z_begin - R, zp_last = z_end - 1 + R;
        const int zlo = max(zp_first, max(0, -g.zoff));
        const int zhi = min(zp_last, min(g.S - 1, g.Sg - 1 - g.zoff));
        const T *src = in + comp * g.N;
        T *dst = out + comp * g.N;
        const int gx = x0 + xh * 32 + lane;  // owned column
        // per-position normalisers (STRICT: weight sums to divide by; otherwise reciprocals, applied once at the end)
        T ny[OY];
#pragma unroll
        for (int j = 0; j < OY; ++j) ny[j] = tab_y[y0 + yl0 + j];
        if constexpr (!STRICT) {
            const T nxv = tab_x[gx];
#pragma unroll
            for (int j = 0; j < OY; ++j) ny[j] *= nxv;
        }
        T win[OY][TAPS];  // z ring of xy-smoothed values, compile-time indices (one copy of the step per ring phase)
#pragma unroll
        for (int j = 0; j < OY; ++j)
#pragma unroll
            for (int kk = 0; kk < TAPS; ++kk) win[j][kk] = T(0);
        T nz_next = tab_z[z_begin];
        const bool col_ok = (gx < g.C);
        bool bad = false;  // some output of this thread came out non-finite
        // running output pointer: row yl0 of the tile, column gx, slice of the next output
        T *outp = dst + static_cast<long long>(z_begin) * plane + static_cast<long long>(y0 + yl0) * g.C + gx;
        const int rows_ok = col_ok ? max(0, min(OY, g.R - (y0 + yl0))) : 0;  // owned rows inside the grid

        auto step = [&](auto ph, const int zp) {
            constexpr int P = decltype(ph)::value;
            const int zo = zp - R;  // output slice completed by this step
            if (zp >= zlo && zp <= zhi) {
                mbar_wait(full_x + b, ph_x);
                const T *sx = reinterpret_cast<const T *>(reinterpret_cast<const unsigned char *>(s_x) + b * Cfg::sx_stride) +
                              yl0 * SXP + xh * 32 + lane;
                // y pass: rows streamed once; output j takes row m as its tap k = m - j (ascending k: the reference's order)
#pragma unroll
                for (int m = 0; m < OY + 2 * R; ++m) {
                    const T t = sx[m * SXP];
#pragma unroll
                    for (int j = 0; j < OY; ++j) {
                        const int kk = m - j;
                        if (kk < 0 || kk >= TAPS) continue;
                        if (kk == 0)
                            win[j][P] = DCMA_WY(0) * t;
                        else
                            win[j][P] += DCMA_WY(kk) * t;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_x + b);  // the tile is in registers: the producers may overwrite it
                if (++b == NXB) {
                    b = 0;
                    ph_x ^= 1;
                }
                if constexpr (STRICT) {
#pragma unroll
                    for (int j = 0; j < OY; ++j) win[j][P] = win[j][P] / ny[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < OY; ++j) win[j][P] = T(0);
            }
            const T nz = nz_next;  // requested one step ago
            nz_next = tab_z[min(max(zo + 1, z_begin), z_end - 1)];
            if (zo < z_begin) return;  // warm-up slices of this segment
#pragma unroll
            for (int j = 0; j < OY; ++j) {
                T zs = DCMA_WZ(0) * win[j][(P + 1) % TAPS];  // tap k of the z pass (oldest first) lives in slot (P+1+k) % TAPS
#pragma unroll
                for (int kk = 1; kk < TAPS; ++kk) zs += DCMA_WZ(kk) * win[j][(P + 1 + kk) % TAPS];
                T o;
                if constexpr (STRICT)
                    o = zs / nz;
                else
                    o = zs * (ny[j] * nz);
                if (j < rows_ok) {
                    // x*0 is 0 for finite x and NaN otherwise.  A non-finite result <=> a non-finite tap somewhere in the
                    // (2r+1)^3 support: such outputs are redone literally after the march (smooth_fix_column)
                    bad |= !(o * T(0) == T(0));
                    outp[static_cast<long long>(j) * g.C] = o;
                }
            }
            outp += plane;
        };
        for (int zb = zp_first; zb <= zp_last; zb += TAPS) unroll_phases<TAPS, 0>(step, zb, zp_last);
        if (bad) smooth_fix_column<T>(src, dst, g, dwx, dwy, dwz, W.r[0], W.r[1], W.r[2], gx, y0 + yl0, OY, z_begin, z_end);
    }
#undef DCMA_WX
#undef DCMA_WY
#undef DCMA_WZ
}

}  // namespace DCMA_NS
