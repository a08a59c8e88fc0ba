// dcma_b200 -- DCMA_B200_FIELD_FP64_STRICT: compiled with -fmad=false; true IEEE divisions; same operation order as
// the reference => the deformation field is bit-identical to the reference's x86-64 (no-FMA) build.
#define DCMA_NS dcma_f64s
#include "engine_impl.cuh"

namespace dcma {
EngineBase *make_engine_f64_strict(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream, const SlabSpec &slab) {
    return new dcma_f64s::Engine<double, true>(g, p, stream, slab);
}

// warp_moving (cc:467-550) on host buffers: out[p] = image(p + D(p)), reference validity rule, optional NaN replacement.
void warp_image_host(const dcma_b200_geom &geom, const float *image, const double *deformation_aos, bool use_oob,
                     float oob_value, float *out, int device) {
    using namespace dcma_f64s;
    if (geom.slices < 1 || geom.rows < 1 || geom.cols < 1 || geom.channels < 1)
        throw std::invalid_argument("volume extents and channel count must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw CudaError(cudaErrorNoDevice, "no CUDA device is available (dcma_b200 has no CPU fallback)");
    if (device >= 0) DCMA_CUDA(cudaSetDevice(device));
    Geo g;
    g.S = static_cast<int>(geom.slices);
    g.Sg = g.S;
    g.zoff = 0;
    g.zb = 0;
    g.ze = g.S;
    g.R = static_cast<int>(geom.rows);
    g.C = static_cast<int>(geom.cols);
    g.CH = static_cast<int>(geom.channels);
    g.N = geom.slices * geom.rows * geom.cols;
    g.px = geom.pxl_dx;
    g.py = geom.pxl_dy;
    g.pz = geom.pxl_dz;
    g.ipx = 1.0 / g.px;
    g.ipy = 1.0 / g.py;
    g.ipz = 1.0 / g.pz;
    g.is2d = (g.S == 1);
    const long long N = g.N;
    float *dM = nullptr, *dW = nullptr;
    double *dA = nullptr, *dS = nullptr;
    auto cleanup = [&]() {
        cudaFree(dM);
        cudaFree(dW);
        cudaFree(dA);
        cudaFree(dS);
    };
    try {
        DCMA_CUDA(cudaMalloc(&dM, sizeof(float) * N * g.CH));
        DCMA_CUDA(cudaMalloc(&dW, sizeof(float) * N * g.CH));
        DCMA_CUDA(cudaMalloc(&dA, sizeof(double) * 3 * N));
        DCMA_CUDA(cudaMalloc(&dS, sizeof(double) * 3 * N));
        DCMA_CUDA(cudaMemcpy(dM, image, sizeof(float) * N * g.CH, cudaMemcpyHostToDevice));
        DCMA_CUDA(cudaMemcpy(dA, deformation_aos, sizeof(double) * 3 * N, cudaMemcpyHostToDevice));
        const int ntx = (g.C + kTileX - 1) / kTileX, nty = (g.R + kTileY - 1) / kTileY;
        if (nty > 65535) throw std::invalid_argument("too many rows");
        const int zlen = 16;
        const dim3 grid(ntx, nty, (g.S + zlen - 1) / zlen);
        const int grid_lin = static_cast<int>(std::min<long long>((3 * N + 255) / 256, 148LL * 16));
        k_import_aos<double><<<grid_lin, 256>>>(dA, dS, N, 0, N);
        if (g.CH == 1)
            k_warp<double, true, long long, true><<<grid, dim3(kTileX, kTileY)>>>(dM, dS, dW, g, zlen, use_oob ? 1 : 0, oob_value);
        else
            k_warp<double, true, long long, false><<<grid, dim3(kTileX, kTileY)>>>(dM, dS, dW, g, zlen, use_oob ? 1 : 0, oob_value);
        DCMA_CUDA(cudaGetLastError());
        DCMA_CUDA(cudaMemcpy(out, dW, sizeof(float) * N * g.CH, cudaMemcpyDeviceToHost));
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}
}  // namespace dcma
