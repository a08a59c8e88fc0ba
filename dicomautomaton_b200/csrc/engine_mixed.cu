// dcma_b200 -- DCMA_B200_FIELD_MIXED: the deformation field D is stored in double and every coordinate, interpolation and
// force computation is fp64 exactly as in DCMA_B200_FIELD_FP64; only the UPDATE field U is stored as float and smoothed
// with the fp32 kernel (packed FFMA2).  176 algorithmic B/voxel-update instead of 256.  What makes fp32 *fields* leave the
// tolerance at full size is the storage of D (an ulp of D moves the sample point of ill-conditioned voxels, DESIGN.md
// section 2); U is a bounded per-iteration increment (|U| <= max_update_magnitude voxels) whose float rounding, ~1e-7
// voxel, is added once and smoothed -- the reference itself narrows half of the gathered U values to float (cc:439-442).
#define DCMA_NS dcma_mixed
#include "engine_impl.cuh"

namespace dcma {
EngineBase *make_engine_mixed(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream, const SlabSpec &slab) {
    return new dcma_mixed::Engine<double, false, float>(g, p, stream, slab);
}
}  // namespace dcma
