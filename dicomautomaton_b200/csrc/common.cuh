// dcma_b200 -- shared device/host declarations for the Demons kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace dcma {

struct CudaError : public std::runtime_error {
    cudaError_t code;
    CudaError(cudaError_t c, const std::string &what) : std::runtime_error(what), code(c) {}
};

#define DCMA_CUDA(expr)                                                                                      \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) {                                                                             \
            throw ::dcma::CudaError(_e, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" +   \
                                            __FILE__ + ":" + std::to_string(__LINE__) + ")");               \
        }                                                                                                    \
    } while (0)

// Geometry of the dense volume: header of `demons_volume<T>` (reference src/Alignment_Demons.h:64-74).
// A z-slab engine holds only the slices [zoff, zoff + S) of a volume of Sg slices (its owned slices plus halo planes
// that neighbouring slabs fill in) and computes the local slices [zb, ze); a whole-volume engine has Sg == S, zoff == 0,
// [zb, ze) == [0, S).  Every in-grid / border rule of the reference is evaluated on the GLOBAL slice index z + zoff.
struct Geo {
    int S, R, C, CH;     // LOCAL slices held in the buffers, rows, cols, channels
    int Sg, zoff;        // slices of the whole volume; global index of local slice 0
    int zb, ze;          // local slice range this engine computes (owned slices)
    long long N;         // S*R*C local voxels (component stride of the SoA fields)
    double px, py, pz;   // spacing in mm (x=column, y=row, z=slice)
    double ipx, ipy, ipz;  // reciprocals, used only by the non-strict modes
    int is2d;            // Sg == 1: no z interpolation (cc:383, cc:477)
};

// Device-resident loop state: replaces the host variables of AlignViaDemons' loop (reference cc:980-995).
struct Ctrl {
    double prev_mse;         // +inf before the first iteration (cc:980)
    double last_mse;         // MSE of the most recent FORCE stage
    long long last_n;        // valid-voxel count of the most recent FORCE stage
    long long iters_done;    // compute_single_iteration equivalents executed since load
    long long converged_at;  // loop iteration index at which the stop rule fired, LLONG_MAX otherwise
    double threshold;        // convergence_threshold
    unsigned int ticket;     // blocks of the running force kernel that have published their MSE partial
    unsigned int halo_overflow;  // slab engines: a compose gather needed a plane outside the local buffers
    double last_sum;         // sum of squared differences / valid count of the most recent FORCE stage (this slab)
};

constexpr long long kNotConverged = 0x7fffffffffffffffLL;

}  // namespace dcma
