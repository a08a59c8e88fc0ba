// dcma_b200 -- histogram matching of the moving image to the fixed image on the GPU:
// reference AlignViaDemonsHelpers::histogram_match, src/Alignment_Demons.cc:760-891, restated step by step with the
// SAME double arithmetic so that the result is bit-identical to the reference's:
//   1. percentile bounds: sorted_finite[size_t(p * (n - 1))]  (cc:797-809) -- here an exact order statistic by a
//      4-pass radix SELECT over the float keys (no sort of N values);
//   2. per-image histograms of the values inside [min, max], bin = min(bins-1, int64((v-min)/(max-min)*bins)) (cc:821-835)
//      -- integer counts (the reference adds 1.0 to doubles: exact), shared-memory privatised;
//   3. running sums, normalisation, first-bin->=quantile lookup table (cc:838-869) -- tiny, on the host, same order;
//   4. per-voxel mapping with clamping outside [src_min, src_max] (cc:872-888).
// Returns the source unchanged when there are no finite values or a range is degenerate (cc:791-794, 812-815).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "common.cuh"
#include "engine.h"

namespace dcma {
namespace {

// float -> unsigned key with the same order as the float comparison (finite values; -0 < +0 is harmless: equal as doubles)
__device__ __forceinline__ unsigned sortable(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float unsortable(unsigned k) {
    const unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
#ifdef __CUDA_ARCH__
    f = __uint_as_float(b);
#else
    std::memcpy(&f, &b, sizeof(f));
#endif
    return f;
}
__device__ __forceinline__ bool finite_f32(float v) { return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u; }

// histogram of the 8-bit digit at `shift` over the finite values whose higher digits equal `prefix` (mask selects them)
__global__ void __launch_bounds__(256)
    k_digit_hist(const float *__restrict__ v, const long long n, const unsigned prefix, const unsigned mask, const int shift,
                 unsigned long long *__restrict__ hist /* 256 */, unsigned long long *__restrict__ n_finite) {
    __shared__ unsigned long long s[256];
    s[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long fin = 0;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float x = v[i];
        if (!finite_f32(x)) continue;
        ++fin;
        const unsigned k = sortable(x);
        if ((k & mask) == prefix) atomicAdd(&s[(k >> shift) & 0xffu], 1ull);
    }
    __syncthreads();
    if (s[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s[threadIdx.x]);
    if (n_finite && fin) atomicAdd(n_finite, fin);
}

constexpr int kMaxSharedBins = 4096;

__global__ void __launch_bounds__(256)
    k_range_hist(const float *__restrict__ v, const long long n, const double lo, const double hi, const long long bins,
                 unsigned long long *__restrict__ hist, const int use_shared) {
    extern __shared__ unsigned long long s_bins[];
    if (use_shared) {
        for (long long b = threadIdx.x; b < bins; b += blockDim.x) s_bins[b] = 0;
        __syncthreads();
    }
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float x = v[i];
        if (!finite_f32(x)) continue;
        const double val = static_cast<double>(x);
        if (val >= lo && val <= hi) {
            // (val - min) / (max - min) * bins, IEEE division then multiplication, truncated: cc:823-824
            long long b = static_cast<long long>(__dmul_rn(__ddiv_rn(__dsub_rn(val, lo), __dsub_rn(hi, lo)), static_cast<double>(bins)));
            b = min(bins - 1, b);
            if (use_shared)
                atomicAdd(&s_bins[b], 1ull);
            else
                atomicAdd(&hist[b], 1ull);
        }
    }
    if (use_shared) {
        __syncthreads();
        for (long long b = threadIdx.x; b < bins; b += blockDim.x)
            if (s_bins[b]) atomicAdd(&hist[b], s_bins[b]);
    }
}

__global__ void __launch_bounds__(256)
    k_apply_lut(const float *__restrict__ src, float *__restrict__ out, const long long n, const double lo, const double hi,
                const double ref_lo, const double ref_hi, const long long bins, const double *__restrict__ lut) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        float x = src[i];
        if (finite_f32(x)) {  // cc:875-886
            const double val = static_cast<double>(x);
            if (val < lo) {
                x = static_cast<float>(ref_lo);
            } else if (val > hi) {
                x = static_cast<float>(ref_hi);
            } else {
                long long b = static_cast<long long>(__dmul_rn(__ddiv_rn(__dsub_rn(val, lo), __dsub_rn(hi, lo)), static_cast<double>(bins)));
                b = min(bins - 1, b);
                x = static_cast<float>(lut[b]);
            }
        }
        out[i] = x;
    }
}

int grid_for(long long n) { return static_cast<int>(std::max<long long>(1, std::min<long long>((n + 255) / 256, 148LL * 8))); }

// exact k-th smallest (0-based) of the finite values; *n_finite_out = their number (valid after the first pass)
float select_kth(const float *d_v, long long n, unsigned long long *d_hist, unsigned long long *d_nfin, long long k) {
    unsigned prefix = 0, mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        DCMA_CUDA(cudaMemset(d_hist, 0, 256 * sizeof(unsigned long long)));
        k_digit_hist<<<grid_for(n), 256>>>(d_v, n, prefix, mask, shift, d_hist, nullptr);
        DCMA_CUDA(cudaGetLastError());
        unsigned long long h[256];
        DCMA_CUDA(cudaMemcpy(h, d_hist, sizeof(h), cudaMemcpyDeviceToHost));
        int d = 0;
        for (; d < 256; ++d) {
            if (static_cast<unsigned long long>(k) < h[d]) break;
            k -= static_cast<long long>(h[d]);
        }
        if (d == 256) throw std::logic_error("radix select ran past the end");
        prefix |= static_cast<unsigned>(d) << shift;
        mask |= 0xffu << shift;
    }
    (void)d_nfin;
    return unsortable(prefix);
}

long long count_finite(const float *d_v, long long n, unsigned long long *d_hist, unsigned long long *d_nfin) {
    DCMA_CUDA(cudaMemset(d_hist, 0, 256 * sizeof(unsigned long long)));
    DCMA_CUDA(cudaMemset(d_nfin, 0, sizeof(unsigned long long)));
    k_digit_hist<<<grid_for(n), 256>>>(d_v, n, 0u, 0u, 24, d_hist, d_nfin);
    DCMA_CUDA(cudaGetLastError());
    unsigned long long c = 0;
    DCMA_CUDA(cudaMemcpy(&c, d_nfin, sizeof(c), cudaMemcpyDeviceToHost));
    return static_cast<long long>(c);
}

std::vector<double> cdf_of(const float *d_v, long long n, double lo, double hi, long long bins, unsigned long long *d_bins) {
    DCMA_CUDA(cudaMemset(d_bins, 0, sizeof(unsigned long long) * bins));
    const int use_shared = bins <= kMaxSharedBins;
    k_range_hist<<<grid_for(n), 256, use_shared ? sizeof(unsigned long long) * bins : 0>>>(d_v, n, lo, hi, bins, d_bins, use_shared);
    DCMA_CUDA(cudaGetLastError());
    std::vector<unsigned long long> h(static_cast<size_t>(bins));
    DCMA_CUDA(cudaMemcpy(h.data(), d_bins, sizeof(unsigned long long) * bins, cudaMemcpyDeviceToHost));
    std::vector<double> cdf(static_cast<size_t>(bins));
    for (long long i = 0; i < bins; ++i) cdf[i] = static_cast<double>(h[i]);  // the reference counts in doubles (exact)
    for (long long i = 1; i < bins; ++i) cdf[i] += cdf[i - 1];                // cc:838-841
    const double total = cdf.back();
    if (total > 0.0)
        for (auto &c : cdf) c /= total;  // cc:846-851
    return cdf;
}

}  // namespace

// returns 1 when the source was mapped, 0 when it is returned unchanged (no finite values / degenerate range)
int histogram_match_host(const float *source, long long n_src, const float *reference, long long n_ref, long long num_bins,
                         double outlier_fraction, float *out, int device) {
    if (n_src < 1 || n_ref < 1) throw std::invalid_argument("Cannot perform histogram matching: image collection is empty");
    if (num_bins < 1) throw std::invalid_argument("histogram_bins must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw CudaError(cudaErrorNoDevice, "no CUDA device is available (dcma_b200 has no CPU fallback)");
    if (device >= 0) DCMA_CUDA(cudaSetDevice(device));
    float *dS = nullptr, *dR = nullptr;
    unsigned long long *d_hist = nullptr, *d_nfin = nullptr, *d_bins = nullptr;
    double *d_lut = nullptr;
    auto cleanup = [&]() {
        cudaFree(dS);
        cudaFree(dR);
        cudaFree(d_hist);
        cudaFree(d_nfin);
        cudaFree(d_bins);
        cudaFree(d_lut);
    };
    int mapped = 0;
    try {
        DCMA_CUDA(cudaMalloc(&dS, sizeof(float) * n_src));
        DCMA_CUDA(cudaMalloc(&dR, sizeof(float) * n_ref));
        DCMA_CUDA(cudaMalloc(&d_hist, 256 * sizeof(unsigned long long)));
        DCMA_CUDA(cudaMalloc(&d_nfin, sizeof(unsigned long long)));
        DCMA_CUDA(cudaMalloc(&d_bins, sizeof(unsigned long long) * num_bins));
        DCMA_CUDA(cudaMalloc(&d_lut, sizeof(double) * num_bins));
        DCMA_CUDA(cudaMemcpy(dS, source, sizeof(float) * n_src, cudaMemcpyHostToDevice));
        DCMA_CUDA(cudaMemcpy(dR, reference, sizeof(float) * n_ref, cudaMemcpyHostToDevice));
        const long long ns = count_finite(dS, n_src, d_hist, d_nfin), nr = count_finite(dR, n_ref, d_hist, d_nfin);
        bool unchanged = (ns == 0 || nr == 0);  // cc:791-794
        double src_min = 0, src_max = 0, ref_min = 0, ref_max = 0;
        if (!unchanged) {
            auto idx = [](double p, long long n) { return static_cast<long long>(static_cast<size_t>(p * static_cast<double>(n - 1))); };  // cc:801-804
            src_min = select_kth(dS, n_src, d_hist, d_nfin, idx(outlier_fraction, ns));
            src_max = select_kth(dS, n_src, d_hist, d_nfin, idx(1.0 - outlier_fraction, ns));
            ref_min = select_kth(dR, n_ref, d_hist, d_nfin, idx(outlier_fraction, nr));
            ref_max = select_kth(dR, n_ref, d_hist, d_nfin, idx(1.0 - outlier_fraction, nr));
            unchanged = (src_max <= src_min || ref_max <= ref_min);  // cc:812-815
        }
        if (unchanged) {
            if (out != source) std::memcpy(out, source, sizeof(float) * n_src);
        } else {
            const std::vector<double> src_cdf = cdf_of(dS, n_src, src_min, src_max, num_bins, d_bins);
            const std::vector<double> ref_cdf = cdf_of(dR, n_ref, ref_min, ref_max, num_bins, d_bins);
            std::vector<double> lut(static_cast<size_t>(num_bins));
            for (long long sb = 0; sb < num_bins; ++sb) {  // cc:854-869
                long long rb = 0;
                for (long long i = 0; i < num_bins; ++i)
                    if (ref_cdf[i] >= src_cdf[sb]) {
                        rb = i;
                        break;
                    }
                lut[sb] = ref_min + (ref_max - ref_min) * static_cast<double>(rb) / static_cast<double>(num_bins);
            }
            DCMA_CUDA(cudaMemcpy(d_lut, lut.data(), sizeof(double) * num_bins, cudaMemcpyHostToDevice));
            // dR is no longer needed: reuse nothing, write the result over a fresh view of dS's size
            float *dO = dR;
            if (n_ref < n_src) {
                cudaFree(dR);
                dR = nullptr;
                DCMA_CUDA(cudaMalloc(&dR, sizeof(float) * n_src));
                dO = dR;
            }
            k_apply_lut<<<grid_for(n_src), 256>>>(dS, dO, n_src, src_min, src_max, ref_min, ref_max, num_bins, d_lut);
            DCMA_CUDA(cudaGetLastError());
            DCMA_CUDA(cudaMemcpy(out, dO, sizeof(float) * n_src, cudaMemcpyDeviceToHost));
            mapped = 1;
        }
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
    return mapped;
}

}  // namespace dcma
