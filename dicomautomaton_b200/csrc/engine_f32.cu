// dcma_b200 -- DCMA_B200_FIELD_FP32: float field storage and fp32 smoothing; coordinates, interpolation and the
// force stay fp64.  Fastest mode (136 algorithmic B/voxel-update instead of 256); not bit-comparable to the reference.
#define DCMA_NS dcma_f32
#include "engine_impl.cuh"

namespace dcma {
EngineBase *make_engine_f32(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream, const SlabSpec &slab) {
    return new dcma_f32::Engine<float, false>(g, p, stream, slab);
}
}  // namespace dcma
