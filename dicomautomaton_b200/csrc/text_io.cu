// dcma_b200 -- the text payload of a stored deformation field on the GPU.
//
// deformation_field::write_to (reference src/Alignment_Field.cc:322-357) prints every double of the field on its own
// line at max_digits10; read_from (:359-436) extracts them again with `is >> val`.  The per-image headers are a few
// lines and stay on the host (the reference's own code, see patches/Alignment_Field.patch); the 3 * voxels numbers per
// image are the work:
//
//   format_doubles_host : doubles -> "%.17g\n" lines.  k_format_lines converts one number per thread (text_format.cuh,
//                         exact), packs the lines of a 1024-number tile contiguously in shared memory (block scan of the
//                         line lengths) and stores the tile; k_scan_tiles turns the tile lengths into byte offsets;
//                         k_gather_tiles moves every tile to its final position.  Variable-length output without a
//                         second conversion pass.
//   parse_doubles_host  : white-space separated decimal literals -> doubles.  k_token_count / k_token_index find the
//                         first byte of every token (a non-space byte after a space byte), k_parse_tokens converts one
//                         token per thread (exact) and reports the first token that is not a number.
//
// Both directions are HBM/PCIe-trivial next to the conversion arithmetic (~25 B of text per number); the kernels are
// integer-ALU bound.
#include <algorithm>
#include <cctype>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "engine.h"
#include "text_format.cuh"

namespace dcma {
namespace {

constexpr int kThreads = 256;
constexpr int kPerThread = 4;
constexpr int kTileValues = kThreads * kPerThread;        // numbers per tile
constexpr int kSlot = txt::kMaxChars + 1;                 // longest line, with its '\n'
constexpr int kTileStride = 25600;                        // bytes reserved per tile in the staging buffer
static_assert(kTileStride >= kTileValues * kSlot && kTileStride % 16 == 0, "tile staging stride");

// block-wide exclusive scan of one int per thread (kThreads threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int *warp_sums, int *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();  // warp_sums may still be read from the previous call
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    int base = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        const int s = warp_sums[w];
        if (w < wid) base += s;
        sum += s;
    }
    *total = sum;
    return base + inc - v;
}

// ---- doubles -> lines --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_format_lines(const double *__restrict__ values, long long n, char *__restrict__ stage,
                                                           int *__restrict__ tile_len) {
    __shared__ __align__(16) char text[kTileStride];
    __shared__ char slots[kThreads][kSlot + 3];
    __shared__ int warp_sums[kThreads / 32];
    const long long tile = blockIdx.x;
    int running = 0;
    for (int j = 0; j < kPerThread; ++j) {
        const long long i = tile * kTileValues + static_cast<long long>(j) * kThreads + threadIdx.x;
        char *slot = slots[threadIdx.x];
        int len = 0;
        if (i < n) {
            len = txt::format_g17(values[i], slot);
            slot[len++] = '\n';
        }
        int total;
        const int off = block_exclusive_scan(len, warp_sums, &total);
        char *dst = text + running + off;
        for (int c = 0; c < len; ++c) dst[c] = slot[c];
        running += total;
    }
    __syncthreads();
    uint4 *out = reinterpret_cast<uint4 *>(stage + tile * kTileStride);
    const uint4 *src = reinterpret_cast<const uint4 *>(text);
    for (int c = threadIdx.x; c < (running + 15) / 16; c += kThreads) out[c] = src[c];
    if (threadIdx.x == 0) tile_len[tile] = running;
}

// exclusive scan of n ints into 64-bit offsets, one block; total[0] = sum
__global__ void __launch_bounds__(1024) k_scan_tiles(const int *__restrict__ len, long long n, long long *__restrict__ off,
                                                     long long *__restrict__ total) {
    __shared__ long long part[1024];
    const long long per = (n + 1023) / 1024;
    const long long b = min(n, per * threadIdx.x), e = min(n, b + per);
    long long s = 0;
    for (long long i = b; i < e; ++i) s += len[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int t = 0; t < 1024; ++t) {
            const long long v = part[t];
            part[t] = run;
            run += v;
        }
        total[0] = run;
    }
    __syncthreads();
    long long run = part[threadIdx.x];
    for (long long i = b; i < e; ++i) {
        off[i] = run;
        run += len[i];
    }
}

__global__ void __launch_bounds__(kThreads) k_gather_tiles(const char *__restrict__ stage, const int *__restrict__ tile_len,
                                                           const long long *__restrict__ tile_off, char *__restrict__ out) {
    const long long tile = blockIdx.x;
    const int len = tile_len[tile];
    const char *src = stage + tile * kTileStride;
    char *dst = out + tile_off[tile];
    for (int c = threadIdx.x; c < len; c += kThreads) dst[c] = src[c];
}

// ---- lines -> doubles --------------------------------------------------------------------------------------------------
constexpr int kBytesPerThread = 16;
constexpr int kTileBytes = kThreads * kBytesPerThread;
constexpr int kMaxToken = 1024;

__device__ __forceinline__ bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }  // isspace, "C" locale

// token starts among the bytes [b, e) of this thread; bit i of the result = byte b + i starts a token
__device__ __forceinline__ unsigned token_start_mask(const char *__restrict__ text, long long b, long long e) {
    unsigned mask = 0;
    bool prev_space = b == 0 ? true : is_space(text[b - 1]);
    for (long long i = b; i < e; ++i) {
        const bool sp = is_space(text[i]);
        if (!sp && prev_space) mask |= 1u << static_cast<int>(i - b);
        prev_space = sp;
    }
    return mask;
}

__global__ void __launch_bounds__(kThreads) k_token_count(const char *__restrict__ text, long long nbytes, int *__restrict__ tile_count) {
    __shared__ int warp_sums[kThreads / 32];
    const long long b = min(nbytes, (static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x) * kBytesPerThread);
    const long long e = min(nbytes, b + kBytesPerThread);
    int total;
    block_exclusive_scan(__popc(token_start_mask(text, b, e)), warp_sums, &total);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kThreads) k_token_index(const char *__restrict__ text, long long nbytes,
                                                          const long long *__restrict__ tile_off, long long max_tokens,
                                                          long long *__restrict__ token_pos) {
    __shared__ int warp_sums[kThreads / 32];
    const long long b = min(nbytes, (static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x) * kBytesPerThread);
    const long long e = min(nbytes, b + kBytesPerThread);
    unsigned mask = token_start_mask(text, b, e);
    int total;
    long long idx = tile_off[blockIdx.x] + block_exclusive_scan(__popc(mask), warp_sums, &total);
    while (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        if (idx < max_tokens) token_pos[idx] = b + bit;
        ++idx;
    }
}

struct ParseReport {
    long long first_bad;  // index of the first token that is not a number (LLONG_MAX if none)
    long long last_end;   // byte offset just past token n_tokens - 1
    int bad_status;       // its txt::ParseStatus
};

__global__ void __launch_bounds__(kThreads) k_parse_tokens(const char *__restrict__ text, long long nbytes,
                                                           const long long *__restrict__ token_pos, long long n_tokens,
                                                           double *__restrict__ out, ParseReport *__restrict__ report,
                                                           int *__restrict__ status_of) {
    const long long j = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x;
    if (j >= n_tokens) return;
    const long long b = token_pos[j];
    long long e = b;
    while (e < nbytes && e - b <= kMaxToken && !is_space(text[e])) ++e;
    double v = 0.0;
    const int st = (e - b > kMaxToken) ? static_cast<int>(txt::kParseMalformed) : txt::parse_double(text + b, static_cast<int>(e - b), &v);
    out[j] = v;
    if (st != txt::kParseOk) {
        status_of[j] = st;
        atomicMin(reinterpret_cast<unsigned long long *>(&report->first_bad), static_cast<unsigned long long>(j));
    }
    if (j == n_tokens - 1) report->last_end = e;
}

struct DeviceBuffers {
    std::vector<void *> ptrs;
    ~DeviceBuffers() {
        for (void *p : ptrs) cudaFree(p);
    }
    template <class T>
    T *alloc(size_t count) {
        void *p = nullptr;
        DCMA_CUDA(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
        ptrs.push_back(p);
        return static_cast<T *>(p);
    }
};

void select_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw CudaError(cudaErrorNoDevice, "no CUDA device is available (dcma_b200 has no CPU fallback)");
    if (device >= 0) DCMA_CUDA(cudaSetDevice(device));
}

}  // namespace

// values[0..n) -> out: each number as printf("%.17g\n").  Returns the number of bytes written; throws
// std::invalid_argument when `capacity` is too small (25 * n always suffices).
long long format_doubles_host(const double *values, long long n, char *out, long long capacity, int device) {
    if (n < 0 || capacity < 0) throw std::invalid_argument("negative size");
    if (n == 0) return 0;
    select_device(device);
    const long long chunk = std::min<long long>(n, 8ll << 20);  // numbers per pass: 64 MiB in, <= 200 MiB of text
    const long long max_tiles = (chunk + kTileValues - 1) / kTileValues;
    DeviceBuffers dev;
    double *d_values = dev.alloc<double>(chunk);
    char *d_stage = dev.alloc<char>(max_tiles * kTileStride);
    char *d_text = dev.alloc<char>(max_tiles * kTileStride);
    int *d_len = dev.alloc<int>(max_tiles);
    long long *d_off = dev.alloc<long long>(max_tiles);
    long long *d_total = dev.alloc<long long>(1);
    long long written = 0;
    for (long long first = 0; first < n; first += chunk) {
        const long long m = std::min(chunk, n - first);
        const long long tiles = (m + kTileValues - 1) / kTileValues;
        DCMA_CUDA(cudaMemcpy(d_values, values + first, sizeof(double) * m, cudaMemcpyHostToDevice));
        k_format_lines<<<static_cast<unsigned>(tiles), kThreads>>>(d_values, m, d_stage, d_len);
        DCMA_CUDA(cudaGetLastError());
        k_scan_tiles<<<1, 1024>>>(d_len, tiles, d_off, d_total);
        DCMA_CUDA(cudaGetLastError());
        k_gather_tiles<<<static_cast<unsigned>(tiles), kThreads>>>(d_stage, d_len, d_off, d_text);
        DCMA_CUDA(cudaGetLastError());
        long long bytes = 0;
        DCMA_CUDA(cudaMemcpy(&bytes, d_total, sizeof bytes, cudaMemcpyDeviceToHost));
        if (written + bytes > capacity)
            throw std::invalid_argument("text buffer too small: need more than " + std::to_string(capacity) + " bytes (25 per number always suffices)");
        DCMA_CUDA(cudaMemcpy(out + written, d_text, static_cast<size_t>(bytes), cudaMemcpyDeviceToHost));
        written += bytes;
    }
    return written;
}

// Reads exactly n white-space separated numbers from text[0..nbytes) into out, like n stream extractions `is >> val`.
// Returns the number of bytes consumed (up to and including the last character of the n-th number).  Throws
// std::invalid_argument when the text holds fewer than n numbers or a token is not a finite decimal number.
long long parse_doubles_host(const char *text, long long nbytes, long long n, double *out, int device) {
    if (n < 0 || nbytes < 0) throw std::invalid_argument("negative size");
    if (n == 0) return 0;
    select_device(device);
    const long long chunk_bytes = std::min<long long>(std::max<long long>(nbytes, 1), 192ll << 20);
    const long long max_tiles = (chunk_bytes + kTileBytes - 1) / kTileBytes;
    const long long max_tokens = std::min<long long>(n, (chunk_bytes + 1) / 2);
    DeviceBuffers dev;
    char *d_text = dev.alloc<char>(chunk_bytes);
    int *d_count = dev.alloc<int>(max_tiles);
    long long *d_off = dev.alloc<long long>(max_tiles);
    long long *d_total = dev.alloc<long long>(1);
    long long *d_pos = dev.alloc<long long>(max_tokens);
    double *d_out = dev.alloc<double>(max_tokens);
    int *d_status = dev.alloc<int>(max_tokens);
    ParseReport *d_report = dev.alloc<ParseReport>(1);
    long long done = 0, base = 0, consumed = 0;
    while (done < n) {
        if (base >= nbytes) throw std::invalid_argument("Pixel data could not be read: the text ends after " + std::to_string(done) + " of " + std::to_string(n) + " numbers");
        long long m = std::min(chunk_bytes, nbytes - base);
        if (base + m < nbytes) {  // do not cut a token: end the chunk after its last white-space byte
            long long cut = m;
            while (cut > 0 && !std::isspace(static_cast<unsigned char>(text[base + cut - 1]))) --cut;
            if (cut == 0) throw std::invalid_argument("Pixel data could not be read: a token is longer than the read buffer");
            m = cut;
        }
        const long long tiles = (m + kTileBytes - 1) / kTileBytes;
        DCMA_CUDA(cudaMemcpy(d_text, text + base, static_cast<size_t>(m), cudaMemcpyHostToDevice));
        k_token_count<<<static_cast<unsigned>(tiles), kThreads>>>(d_text, m, d_count);
        DCMA_CUDA(cudaGetLastError());
        k_scan_tiles<<<1, 1024>>>(d_count, tiles, d_off, d_total);
        DCMA_CUDA(cudaGetLastError());
        long long found = 0;
        DCMA_CUDA(cudaMemcpy(&found, d_total, sizeof found, cudaMemcpyDeviceToHost));
        const long long take = std::min(found, n - done);
        if (take > 0) {
            k_token_index<<<static_cast<unsigned>(tiles), kThreads>>>(d_text, m, d_off, take, d_pos);
            DCMA_CUDA(cudaGetLastError());
            const ParseReport init = {kNotConverged, 0, 0};
            DCMA_CUDA(cudaMemcpy(d_report, &init, sizeof init, cudaMemcpyHostToDevice));
            k_parse_tokens<<<static_cast<unsigned>((take + kThreads - 1) / kThreads), kThreads>>>(d_text, m, d_pos, take, d_out, d_report,
                                                                                                 d_status);
            DCMA_CUDA(cudaGetLastError());
            ParseReport rep;
            DCMA_CUDA(cudaMemcpy(&rep, d_report, sizeof rep, cudaMemcpyDeviceToHost));
            if (rep.first_bad != kNotConverged) {
                int st = 0;
                long long pos = 0;
                DCMA_CUDA(cudaMemcpy(&st, d_status + rep.first_bad, sizeof st, cudaMemcpyDeviceToHost));
                DCMA_CUDA(cudaMemcpy(&pos, d_pos + rep.first_bad, sizeof pos, cudaMemcpyDeviceToHost));
                const char *why = st == txt::kParseOverflow ? "is out of the range of a double"
                                  : st == txt::kParseUndecided ? "has too many digits to be rounded exactly"
                                                               : "is not a decimal number";
                throw std::invalid_argument("Pixel data could not be read: number " + std::to_string(done + rep.first_bad) + " (byte " +
                                            std::to_string(base + pos) + ") " + why);
            }
            DCMA_CUDA(cudaMemcpy(out + done, d_out, sizeof(double) * take, cudaMemcpyDeviceToHost));
            done += take;
            consumed = base + rep.last_end;
        }
        base += m;
    }
    return consumed;
}

}  // namespace dcma
