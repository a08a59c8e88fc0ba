/* Plain-C caller of the dcma_b200 boundary: what a maintainer's binding does, without any C++ or Python.
 *
 *   gcc -std=c99 -I include examples/register_from_c.c -L dicomautomaton_b200/lib -ldcma_b200 \
 *       -Wl,-rpath,$PWD/dicomautomaton_b200/lib -lm -o /tmp/register_from_c && /tmp/register_from_c
 *
 * Registers a small synthetic pair (a Gaussian blob shifted by 1.5 voxels along x) with the diffeomorphic variant,
 * warps the moving image with the result and prints the MSE before and after.  Needs a B200: without a device every
 * compute entry point returns DCMA_B200_ENODEVICE (there is no CPU fallback) and the program says so.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "dcma_b200/demons_c_api.h"

static float blob(double x, double y, double z) {
    const double r2 = (x - 24.0) * (x - 24.0) + (y - 20.0) * (y - 20.0) + (z - 12.0) * (z - 12.0);
    return (float)(1000.0 * exp(-r2 / (2.0 * 36.0)) + 5.0 * sin(0.9 * x + 0.3 * y) + 5.0 * sin(0.7 * y + 0.4 * z));
}

int main(void) {
    const dcma_b200_geom g = {24, 40, 48, 1, 1.0, 1.0, 1.0};
    const size_t n = (size_t)(g.slices * g.rows * g.cols);
    float *fixed = malloc(n * sizeof(float)), *moving = malloc(n * sizeof(float)), *warped = malloc(n * sizeof(float));
    double *field = malloc(3 * n * sizeof(double)), *history = malloc(60 * sizeof(double));
    char err[512] = {0};
    if (!fixed || !moving || !warped || !field || !history) return 2;
    for (int64_t z = 0; z < g.slices; ++z)
        for (int64_t y = 0; y < g.rows; ++y)
            for (int64_t x = 0; x < g.cols; ++x) {
                const size_t i = (size_t)((z * g.rows + y) * g.cols + x);
                fixed[i] = blob((double)x, (double)y, (double)z);
                moving[i] = blob((double)x - 1.5, (double)y, (double)z);
            }
    printf("dcma_b200 ABI %d, %d CUDA device(s)\n", dcma_b200_abi_version(), dcma_b200_device_count());

    dcma_b200_demons_params p;
    dcma_b200_demons_params_default(&p); /* the reference's defaults (src/Alignment_Demons.h:20-61) */
    p.max_iterations = 60;
    p.convergence_threshold = 0.0;
    p.use_diffeomorphic = 1;
    p.deformation_field_smoothing_sigma = 1.5;
    p.update_field_smoothing_sigma = 1.5;
    p.verbosity = 0;
    dcma_b200_run_stats st;
    int rc = dcma_b200_demons_register(&g, fixed, moving, &p, field, history, &st, err, sizeof err);
    if (rc != DCMA_B200_OK) {
        printf("registration failed (%d): %s\n", rc, err);
        return rc == DCMA_B200_ENODEVICE ? 0 : 1;
    }
    rc = dcma_b200_warp_image(&g, moving, field, 0.0f, 0, warped, -1, err, sizeof err);
    if (rc != DCMA_B200_OK) {
        printf("warp failed (%d): %s\n", rc, err);
        return 1;
    }
    double before = 0.0, after = 0.0;
    size_t cnt = 0;
    for (size_t i = 0; i < n; ++i) {
        if (warped[i] != warped[i]) continue; /* NaN on the faces the field points out of (cc:516-522) */
        before += ((double)fixed[i] - moving[i]) * ((double)fixed[i] - moving[i]);
        after += ((double)fixed[i] - warped[i]) * ((double)fixed[i] - warped[i]);
        ++cnt;
    }
    printf("%lld iterations in %.2f ms (%lld kernel launches): MSE %.3f -> %.3f over %zu voxels; first/last loop MSE %.3f / %.3f\n",
           (long long)st.iterations_done, st.loop_ms, (long long)st.kernel_launches, before / cnt, after / cnt, cnt,
           history[0], history[st.iterations_done - 1]);
    free(fixed);
    free(moving);
    free(warped);
    free(field);
    free(history);
    return after < before ? 0 : 1;
}
