// TEST stand-in for the C-ABI entry points host/Alignment_Demons.cc calls, so that the host side of the drop-in
// AlignViaDemons -- marshalling of both stacks, the grids it describes, and above all the copy of the field into the
// output stack CHUNK BY CHUNK from the notification callback -- can be checked on a machine without a GPU.  The "field"
// is a known function of the voxel index, delivered in chunks that do not line up with slices; the stand-in also
// records what it was given (grids, a checksum of the volumes) for the driver to inspect.  Never part of the product.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "dcma_b200/demons_c_api.h"

extern "C" {
dcma_b200_grid g_fake_fixed_grid, g_fake_moving_grid;
double g_fake_fixed_sum = 0.0, g_fake_moving_sum = 0.0;
int g_fake_calls_stacks = 0, g_fake_calls_plain = 0;
int64_t g_fake_chunk = 1000;  // voxels per notification: deliberately not a multiple of a slice

double fake_field_value(int64_t voxel, int c) { return 0.001 * static_cast<double>(voxel) + 0.25 * c - 3.0; }

int dcma_b200_device_count(void) { return 1; }
int dcma_b200_host_alloc(void **out, size_t bytes, char *, size_t) { *out = std::malloc(bytes ? bytes : 1); return *out ? 0 : -3; }
void dcma_b200_host_free(void *p) { std::free(p); }
void dcma_b200_demons_params_default(dcma_b200_demons_params *p) {
    std::memset(p, 0, sizeof(*p));
    p->max_iterations = 100;
    p->pyramid_levels = 1;
    p->device = -1;
    p->fused = 1;
}
int dcma_b200_demons_register_stacks(const dcma_b200_grid *fg, const float *fixed, const dcma_b200_grid *mg, const float *moving,
                                     const dcma_b200_demons_params *, double *out, double *, dcma_b200_run_stats *stats,
                                     dcma_b200_chunk_fn cb, void *user, char *, size_t) {
    ++g_fake_calls_stacks;
    g_fake_fixed_grid = *fg;
    g_fake_moving_grid = *mg;
    const int64_t nf = fg->geom.slices * fg->geom.rows * fg->geom.cols, nm = mg->geom.slices * mg->geom.rows * mg->geom.cols;
    g_fake_fixed_sum = g_fake_moving_sum = 0.0;
    for (int64_t i = 0; i < nf * fg->geom.channels; ++i) g_fake_fixed_sum += fixed[i];
    for (int64_t i = 0; i < nm * mg->geom.channels; ++i) g_fake_moving_sum += moving[i];
    if (stats) std::memset(stats, 0, sizeof(*stats));
    for (int64_t v0 = 0; v0 < nf; v0 += g_fake_chunk) {  // a chunk is written, THEN announced; later chunks are still garbage
        const int64_t v1 = v0 + g_fake_chunk < nf ? v0 + g_fake_chunk : nf;
        for (int64_t v = v0; v < v1; ++v)
            for (int c = 0; c < 3; ++c) out[3 * v + c] = fake_field_value(v, c);
        if (cb) cb(user, v0, v1);
    }
    return 0;
}
int dcma_b200_demons_register(const dcma_b200_geom *g, const float *, const float *, const dcma_b200_demons_params *, double *out, double *,
                              dcma_b200_run_stats *stats, char *, size_t) {
    ++g_fake_calls_plain;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    const int64_t n = g->slices * g->rows * g->cols;
    for (int64_t v = 0; v < n; ++v)
        for (int c = 0; c < 3; ++c) out[3 * v + c] = fake_field_value(v, c);
    return 0;
}
int dcma_b200_resample_to_grid(const dcma_b200_grid *, const float *image, const dcma_b200_grid *dst, float, float *out, int, char *, size_t) {
    const int64_t n = dst->geom.slices * dst->geom.rows * dst->geom.cols * dst->geom.channels;
    std::memcpy(out, image, sizeof(float) * static_cast<size_t>(n));  // the driver only uses identical grids on this path
    return 0;
}
int dcma_b200_histogram_match(const float *src, int64_t n, const float *, int64_t, int64_t, double, float *out, int32_t *mapped, int, char *,
                              size_t) {
    std::memcpy(out, src, sizeof(float) * static_cast<size_t>(n));
    if (mapped) *mapped = 1;
    return 0;
}
int dcma_b200_warp_image_grid(const dcma_b200_grid *, const double *, const dcma_b200_grid *ig, const float *image, const unsigned char *, float,
                              float *out, int, char *, size_t) {
    const int64_t n = ig->geom.slices * ig->geom.rows * ig->geom.cols * ig->geom.channels;
    std::memcpy(out, image, sizeof(float) * static_cast<size_t>(n));
    return 0;
}
}
