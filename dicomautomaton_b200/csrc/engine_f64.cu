// dcma_b200 -- DCMA_B200_FIELD_FP64: double field storage, double arithmetic, FMA contraction allowed.
#define DCMA_NS dcma_f64
#include "engine_impl.cuh"

namespace dcma {
EngineBase *make_engine_f64(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream, const SlabSpec &slab) {
    return new dcma_f64::Engine<double, false>(g, p, stream, slab);
}
}  // namespace dcma
