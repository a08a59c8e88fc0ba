// Microbenchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a, per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ffma2 tools/ubench/ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2 &d, const float2 a, const float2 b) {
    unsigned long long ra = *reinterpret_cast<const unsigned long long *>(&a), rb = *reinterpret_cast<const unsigned long long *>(&b),
                       rd = *reinterpret_cast<unsigned long long *>(&d);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rd) : "l"(ra), "l"(rb));
    d = *reinterpret_cast<float2 *>(&rd);
}

template <int MODE>
__global__ void k(float *out, float w, int iters, long long *cycles) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 ww = make_float2(w, w);
    const float2 xx = make_float2(w * 0.5f, w * 0.25f);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {
                acc[i].x = fmaf(xx.x, ww.x, acc[i].x);
                acc[i].y = fmaf(xx.y, ww.y, acc[i].y);
            } else {
                ffma2(acc[i], xx, ww);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[MODE] = t1 - t0;
}

int main() {
    float *out;
    long long *cyc, h[2];
    cudaMalloc(&out, sizeof(float) * 148 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 2);
    const int iters = 4096;
    for (int warps = 1; warps <= 8; warps *= 2) {  // warps per SM sub-partition
        const int threads = warps * 4 * 32;
        k<0><<<148, threads>>>(out, 1.0001f, iters, cyc);
        k<1><<<148, threads>>>(out, 1.0001f, iters, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        // per sub-partition: warps * iters * 16 scalar FMA-instr-equivalents (mode 0: 16 FFMA; mode 1: 8 FFMA2)
        printf("warps/SMSP=%d  FFMA: %.3f cyc per warp-FFMA   FFMA2: %.3f cyc per warp-FFMA2 (= %.3f per scalar-equivalent)\n", warps,
               (double)h[0] / (warps * (double)iters * 16), (double)h[1] / (warps * (double)iters * 8),
               (double)h[1] / (warps * (double)iters * 16));
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
