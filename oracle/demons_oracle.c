/* TEST INFRASTRUCTURE ONLY (oracle/).  Plain-C CPU restatement of DICOMautomaton's Demons iteration loop.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference/).  It is our own code, not a
 * copy: the reference is SYCL C++ lambdas over USM pointers; this is C99 loops (OpenMP over z) over the same dense
 * layout  idx = ((z*rows + y)*cols + x)*ch + c  (src/Alignment_Demons.h:77-79).
 *
 * Pinned: tests/test_oracle.py requires this restatement to equal the compiled reference engine
 * (oracle/_ref/libdemons_ref.so, built by oracle/build_ref.py from the reference's own sources) BIT-FOR-BIT on the
 * deformation field, update field, warped image and MSE, stage by stage and over whole loops, and
 * tests/golden/*.npz hold reference-generated vectors for machines where /root/reference is absent.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product path (dicomautomaton_b200/csrc) never does; it fails loudly without its CUDA extension.
 *
 * Build:  gcc -std=c11 -O3 -fopenmp -ffp-contract=off -fPIC -shared demons_oracle.c -o libdemons_oracle.so -lm
 *   -ffp-contract=off: no FMA contraction, so the arithmetic order equals the reference's x86-64 build.
 *   -DORC_FIELD_T=float -DORC_ACC_T=float builds the "fp32-storage emulation" used to predict, on CPU, how far the
 *   GPU's fp32-storage mode may drift from the fp64 reference (it is NOT a parity oracle; suffix _f32).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef ORC_FIELD_T
#define ORC_FIELD_T double /* storage type of gradient/update/deformation: reference uses double (cc:132-134) */
#endif
#ifndef ORC_ACC_T
#define ORC_ACC_T double /* accumulation type inside the Gaussian passes: reference uses double (cc:620-621) */
#endif
#ifndef ORC_SUFFIX
#define ORC_SUFFIX
#endif
#define ORC_CAT2(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define ORC_NAME(n) ORC_CAT(n, ORC_SUFFIX)

typedef ORC_FIELD_T field_t;
typedef ORC_ACC_T acc_t;

typedef struct {
    int64_t slices, rows, cols, channels;
    double pxl_dx, pxl_dy, pxl_dz;
} orc_geom;

typedef struct { /* mirrors AlignViaDemonsParams, src/Alignment_Demons.h:20-61 (fields the loop uses) */
    int64_t max_iterations;
    double convergence_threshold;
    double deformation_field_smoothing_sigma;
    double update_field_smoothing_sigma;
    int32_t use_diffeomorphic;
    int32_t _pad;
    double normalization_factor;
    double max_update_magnitude;
} orc_params;

typedef struct {
    orc_geom g;
    orc_params p;
    size_t n_vox;    /* slices*rows*cols            (cc:153) */
    size_t n_scalar; /* n_vox*channels              (cc:154) */
    size_t n_vector; /* n_vox*3                     (cc:155) */
    float *fixed, *moving, *warped;              /* cc:129-131 */
    field_t *gradient, *update, *deformation;    /* cc:132-134 */
    double *mse_term;                            /* cc:135 */
    int64_t *valid;                              /* cc:136 */
} orc_engine;

/* ---- a4: compute_gradient, src/Alignment_Demons.cc:208-272 -------------------------------------------------- */
static void orc_gradient(orc_engine *E) {
    const int64_t S = E->g.slices, R = E->g.rows, C = E->g.cols, CH = E->g.channels;
    const float *F = E->fixed;
#pragma omp parallel for schedule(static)
    for (int64_t z = 0; z < S; ++z)
        for (int64_t y = 0; y < R; ++y)
            for (int64_t x = 0; x < C; ++x) {
#define IN_IDX(zz, yy, xx) ((((zz)*R + (yy)) * C + (xx)) * CH) /* channel 0 only, cc:225-227 */
                double gx = 0.0, gy = 0.0, gz = 0.0;
                if (C > 1) { /* cc:233-243 */
                    const int64_t xl = (x > 0) ? (x - 1) : x;
                    const int64_t xr = (x + 1 < C) ? (x + 1) : x;
                    const float vl = F[IN_IDX(z, y, xl)], vr = F[IN_IDX(z, y, xr)];
                    if (isfinite(vl) && isfinite(vr) && xr != xl) gx = ((double)vr - (double)vl) / (double)(xr - xl);
                }
                if (R > 1) { /* cc:244-254 */
                    const int64_t yu = (y > 0) ? (y - 1) : y;
                    const int64_t yd = (y + 1 < R) ? (y + 1) : y;
                    const float vu = F[IN_IDX(z, yu, x)], vd = F[IN_IDX(z, yd, x)];
                    if (isfinite(vu) && isfinite(vd) && yd != yu) gy = ((double)vd - (double)vu) / (double)(yd - yu);
                }
                if (S > 1) { /* cc:255-265 */
                    const int64_t zp = (z > 0) ? (z - 1) : z;
                    const int64_t zn = (z + 1 < S) ? (z + 1) : z;
                    const float vp = F[IN_IDX(zp, y, x)], vn = F[IN_IDX(zn, y, x)];
                    if (isfinite(vp) && isfinite(vn) && zn != zp) gz = ((double)vn - (double)vp) / (double)(zn - zp);
                }
#undef IN_IDX
                const size_t o = (size_t)(((z * R + y) * C + x) * 3);
                E->gradient[o + 0] = (field_t)gx;
                E->gradient[o + 1] = (field_t)gy;
                E->gradient[o + 2] = (field_t)gz;
            }
}

/* ---- a5: compute_update_and_mse, src/Alignment_Demons.cc:274-343 -------------------------------------------- */
static void orc_update_and_mse(orc_engine *E, double *mse_out, int64_t *n_out) {
    const double epsilon = 1.0e-10; /* cc:275 */
    const int64_t S = E->g.slices, R = E->g.rows, C = E->g.cols, CH = E->g.channels;
    const double normalization = E->p.normalization_factor, max_update = E->p.max_update_magnitude;
#pragma omp parallel for schedule(static)
    for (int64_t z = 0; z < S; ++z)
        for (int64_t y = 0; y < R; ++y)
            for (int64_t x = 0; x < C; ++x) {
                const size_t v = (size_t)((z * R + y) * C + x);
                const size_t f = v * (size_t)CH, g = v * 3u;
                const float fixed_val = E->fixed[f], moving_val = E->warped[f]; /* channel 0, cc:294-295 */
                if (!(isfinite(fixed_val) && isfinite(moving_val))) {             /* cc:297-304 */
                    E->mse_term[v] = 0.0;
                    E->valid[v] = 0;
                    E->update[g + 0] = 0;
                    E->update[g + 1] = 0;
                    E->update[g + 2] = 0;
                    continue;
                }
                const double diff = (double)fixed_val - (double)moving_val; /* cc:306 */
                E->mse_term[v] = diff * diff;
                E->valid[v] = 1;
                double ux = 0.0, uy = 0.0, uz = 0.0;
                const double gx = (double)E->gradient[g + 0], gy = (double)E->gradient[g + 1],
                             gz = (double)E->gradient[g + 2];
                const double gmag_sq = gx * gx + gy * gy + gz * gz;                        /* cc:314 */
                const double denom = gmag_sq + (diff * diff) / (normalization + epsilon); /* cc:315 */
                if (denom > epsilon) {                                                    /* cc:316-327 */
                    ux = diff * gx / denom;
                    uy = diff * gy / denom;
                    uz = diff * gz / denom;
                    const double mag = sqrt(ux * ux + uy * uy + uz * uz);
                    if (mag > max_update) {
                        const double scale = max_update / mag;
                        ux *= scale;
                        uy *= scale;
                        uz *= scale;
                    }
                }
                E->update[g + 0] = (field_t)ux;
                E->update[g + 1] = (field_t)uy;
                E->update[g + 2] = (field_t)uz;
            }
    /* serial host sum in index order, cc:334-342 */
    double mse = 0.0;
    int64_t n = 0;
    for (size_t i = 0; i < E->n_vox; ++i) {
        mse += E->mse_term[i];
        n += E->valid[i];
    }
    if (n > 0) mse /= (double)n;
    *mse_out = mse;
    *n_out = n;
}

/* ---- a6: add_update, src/Alignment_Demons.cc:345-369 -------------------------------------------------------- */
static void orc_add_update(orc_engine *E) {
    const double px = E->g.pxl_dx, py = E->g.pxl_dy, pz = E->g.pxl_dz;
    const int64_t n = (int64_t)E->n_vox;
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v < n; ++v) {
        const size_t i = (size_t)v * 3u;
        E->deformation[i + 0] = (field_t)((double)E->deformation[i + 0] + (double)E->update[i + 0] * px);
        E->deformation[i + 1] = (field_t)((double)E->deformation[i + 1] + (double)E->update[i + 1] * py);
        E->deformation[i + 2] = (field_t)((double)E->deformation[i + 2] + (double)E->update[i + 2] * pz);
    }
}

/* ---- a7: add_update_diffeomorphic, src/Alignment_Demons.cc:371-465 ------------------------------------------ */
static void orc_add_update_diffeomorphic(orc_engine *E) {
    const int64_t S = E->g.slices, R = E->g.rows, C = E->g.cols;
    const double px = E->g.pxl_dx, py = E->g.pxl_dy, pz = E->g.pxl_dz;
    const int is_2d = (S == 1); /* cc:383 */
    const field_t *U = E->update;
    field_t *D = E->deformation;
#pragma omp parallel for schedule(static)
    for (int64_t z = 0; z < S; ++z)
        for (int64_t y = 0; y < R; ++y)
            for (int64_t x = 0; x < C; ++x) {
                const size_t o = (size_t)(((z * R + y) * C + x) * 3);
                const double dx = (double)D[o + 0], dy = (double)D[o + 1], dz = (double)D[o + 2]; /* cc:397-399 */
                const double sx = (double)x + dx / px;                                            /* cc:403-405 */
                const double sy = (double)y + dy / py;
                const double sz = is_2d ? (double)z : ((double)z + dz / pz);
                const int64_t x0 = (int64_t)floor(sx), y0 = (int64_t)floor(sy), z0 = (int64_t)floor(sz);
                const int64_t x1 = x0 + 1, y1 = y0 + 1, z1 = is_2d ? z0 : (z0 + 1); /* cc:411-416 */
                const double tx = sx - (double)x0, ty = sy - (double)y0;
                const double tz = is_2d ? 0.0 : (sz - (double)z0); /* cc:419-421 */
                double u[3];
                for (int c = 0; c < 3; ++c) {
#define FETCH(zz, yy, xx)                                                                                    \
    (((xx) < 0 || (xx) >= C || (yy) < 0 || (yy) >= R || (zz) < 0 || (zz) >= S)                             \
         ? 0.0                                                                                               \
         : (double)U[(size_t)((((zz)*R + (yy)) * C + (xx)) * 3 + c)]) /* zero outside, cc:424-429 */
                    const double c000 = FETCH(z0, y0, x0), c100 = FETCH(z0, y0, x1);
                    const double c010 = FETCH(z0, y1, x0), c110 = FETCH(z0, y1, x1);
                    /* the four upper-z corners are narrowed to float in the reference, cc:439-442 */
                    const float c001 = (float)(is_2d ? c000 : FETCH(z1, y0, x0));
                    const float c101 = (float)(is_2d ? c100 : FETCH(z1, y0, x1));
                    const float c011 = (float)(is_2d ? c010 : FETCH(z1, y1, x0));
                    const float c111 = (float)(is_2d ? c110 : FETCH(z1, y1, x1));
#undef FETCH
                    const double c00 = c000 * (1.0 - tx) + c100 * tx; /* cc:445-451 */
                    const double c10 = c010 * (1.0 - tx) + c110 * tx;
                    const double c01 = c001 * (1.0 - tx) + c101 * tx;
                    const double c11 = c011 * (1.0 - tx) + c111 * tx;
                    const double c0 = c00 * (1.0 - ty) + c10 * ty;
                    const double c1 = c01 * (1.0 - ty) + c11 * ty;
                    u[c] = c0 * (1.0 - tz) + c1 * tz;
                }
                D[o + 0] = (field_t)((double)D[o + 0] + u[0] * px); /* cc:460-462 */
                D[o + 1] = (field_t)((double)D[o + 1] + u[1] * py);
                D[o + 2] = (field_t)((double)D[o + 2] + u[2] * pz);
            }
}

/* ---- a9: warp_moving, src/Alignment_Demons.cc:467-550 ------------------------------------------------------- */
static void orc_warp_moving(orc_engine *E) {
    const int64_t S = E->g.slices, R = E->g.rows, C = E->g.cols, CH = E->g.channels;
    const double px = E->g.pxl_dx, py = E->g.pxl_dy, pz = E->g.pxl_dz;
    const int is_2d = (S == 1);
    const float oob = NAN; /* cc:468 */
    const field_t *D = E->deformation;
    const float *M = E->moving;
    float *W = E->warped;
#pragma omp parallel for schedule(static)
    for (int64_t z = 0; z < S; ++z)
        for (int64_t y = 0; y < R; ++y)
            for (int64_t x = 0; x < C; ++x) {
                const size_t v = (size_t)((z * R + y) * C + x);
                const double dx = (double)D[v * 3 + 0], dy = (double)D[v * 3 + 1], dz = (double)D[v * 3 + 2];
                if (!(isfinite(dx) && isfinite(dy) && isfinite(dz))) { /* cc:494-500 */
                    for (int64_t c = 0; c < CH; ++c) W[v * (size_t)CH + (size_t)c] = oob;
                    continue;
                }
                const double sx = (double)x + dx / px, sy = (double)y + dy / py; /* cc:501-503 */
                const double sz = is_2d ? (double)z : ((double)z + dz / pz);
                const int64_t x0 = (int64_t)floor(sx), y0 = (int64_t)floor(sy), z0 = (int64_t)floor(sz);
                const int64_t x1 = x0 + 1, y1 = y0 + 1, z1 = is_2d ? z0 : (z0 + 1); /* cc:505-510 */
                const double tx = sx - (double)x0, ty = sy - (double)y0;
                const double tz = is_2d ? 0.0 : (sz - (double)z0);
                for (int64_t c = 0; c < CH; ++c) { /* all channels, cc:515 */
#define SAMPLE(zz, yy, xx)                                                                                   \
    (((xx) < 0 || (xx) >= C || (yy) < 0 || (yy) >= R || (zz) < 0 || (zz) >= S)                             \
         ? oob                                                                                               \
         : M[(size_t)((((zz)*R + (yy)) * C + (xx)) * CH + c)]) /* NaN outside, cc:516-521 */
                    const float c000 = SAMPLE(z0, y0, x0), c100 = SAMPLE(z0, y0, x1);
                    const float c010 = SAMPLE(z0, y1, x0), c110 = SAMPLE(z0, y1, x1);
                    const float c001 = is_2d ? c000 : SAMPLE(z1, y0, x0);
                    const float c101 = is_2d ? c100 : SAMPLE(z1, y0, x1);
                    const float c011 = is_2d ? c010 : SAMPLE(z1, y1, x0);
                    const float c111 = is_2d ? c110 : SAMPLE(z1, y1, x1);
#undef SAMPLE
                    if (!(isfinite(c000) && isfinite(c100) && isfinite(c010) && isfinite(c110) && isfinite(c001) &&
                          isfinite(c101) && isfinite(c011) && isfinite(c111))) { /* cc:535-539 */
                        W[v * (size_t)CH + (size_t)c] = oob;
                        continue;
                    }
                    const double c00 = (double)c000 * (1.0 - tx) + (double)c100 * tx; /* cc:540-546 */
                    const double c10 = (double)c010 * (1.0 - tx) + (double)c110 * tx;
                    const double c01 = (double)c001 * (1.0 - tx) + (double)c101 * tx;
                    const double c11 = (double)c011 * (1.0 - tx) + (double)c111 * tx;
                    const double c0 = c00 * (1.0 - ty) + c10 * ty;
                    const double c1 = c01 * (1.0 - ty) + c11 * ty;
                    W[v * (size_t)CH + (size_t)c] = (float)(c0 * (1.0 - tz) + c1 * tz);
                }
            }
}

/* ---- a8: smooth_on_device, src/Alignment_Demons.cc:563-702 -------------------------------------------------- */
static double *orc_make_kernel(int64_t radius, double sigma_pix) { /* cc:587-597 */
    double *k = (double *)malloc((size_t)(2 * radius + 1) * sizeof(double));
    double sum = 0.0;
    for (int64_t i = -radius; i <= radius; ++i) {
        const double v = exp(-0.5 * ((double)(i * i)) / (sigma_pix * sigma_pix));
        k[i + radius] = v;
        sum += v;
    }
    for (int64_t i = 0; i < 2 * radius + 1; ++i) k[i] /= sum;
    return k;
}

/* one axis pass: out = sum(w*v)/sum(w) over in-bounds finite taps; keeps input if no tap (cc:619-634 etc.) */
static void orc_smooth_pass(const field_t *in, field_t *out, int64_t S, int64_t R, int64_t C, int axis, int64_t rad,
                            const double *k) {
    const int64_t extent = (axis == 0) ? C : (axis == 1) ? R : S;
    const int64_t stride = (axis == 0) ? 3 : (axis == 1) ? C * 3 : R * C * 3;
#pragma omp parallel for schedule(static)
    for (int64_t z = 0; z < S; ++z)
        for (int64_t y = 0; y < R; ++y)
            for (int64_t x = 0; x < C; ++x) {
                const int64_t pos = (axis == 0) ? x : (axis == 1) ? y : z;
                const int64_t base = ((z * R + y) * C + x) * 3;
                for (int64_t c = 0; c < 3; ++c) {
                    acc_t sum = 0, wsum = 0;
                    for (int64_t kk = -rad; kk <= rad; ++kk) {
                        const int64_t q = pos + kk;
                        if (q >= 0 && q < extent) {
                            const acc_t v = (acc_t)in[base + c + kk * stride];
                            if (isfinite(v)) {
                                const acc_t w = (acc_t)k[kk + rad];
                                sum += w * v;
                                wsum += w;
                            }
                        }
                    }
                    out[base + c] = (wsum > 0) ? (field_t)(sum / wsum) : in[base + c];
                }
            }
}

static void orc_smooth(orc_engine *E, field_t *field, double sigma_mm) {
    if (sigma_mm <= 0.0) return; /* cc:564-566 */
    const int64_t S = E->g.slices, R = E->g.rows, C = E->g.cols;
    const double px = E->g.pxl_dx, py = E->g.pxl_dy, pz = E->g.pxl_dz;
    int64_t rad_x = (int64_t)(3.0 * sigma_mm / px), rad_y = (int64_t)(3.0 * sigma_mm / py),
            rad_z = (int64_t)(3.0 * sigma_mm / pz); /* cc:574-576 */
    if (rad_x < 1) rad_x = 1;
    if (rad_y < 1) rad_y = 1;
    if (rad_z < 1) rad_z = 1;
    double *kx = orc_make_kernel(rad_x, sigma_mm / px); /* cc:578-600 */
    double *ky = orc_make_kernel(rad_y, sigma_mm / py);
    double *kz = orc_make_kernel(rad_z, sigma_mm / pz);
    field_t *temp = (field_t *)malloc(E->n_vector * sizeof(field_t));
    orc_smooth_pass(field, temp, S, R, C, 0, rad_x, kx); /* x: field -> temp, cc:609-635 */
    orc_smooth_pass(temp, field, S, R, C, 1, rad_y, ky); /* y: temp -> field, cc:638-664 */
    orc_smooth_pass(field, temp, S, R, C, 2, rad_z, kz); /* z: field -> temp, cc:667-693 */
    memcpy(field, temp, E->n_vector * sizeof(field_t));  /* cc:696 */
    free(temp);
    free(kx);
    free(ky);
    free(kz);
}

/* ---- a3 + a10: engine life cycle and iteration order, src/Alignment_Demons.cc:54-108,152-187 ---------------- */
void *ORC_NAME(orc_engine_create)(const orc_geom *g, const float *fixed, const float *moving, const orc_params *p) {
    orc_engine *E = (orc_engine *)calloc(1, sizeof(orc_engine));
    if (!E) return NULL;
    E->g = *g;
    E->p = *p;
    E->n_vox = (size_t)(g->slices * g->rows * g->cols);
    E->n_scalar = E->n_vox * (size_t)g->channels;
    E->n_vector = E->n_vox * 3u;
    E->fixed = (float *)malloc(E->n_scalar * sizeof(float));
    E->moving = (float *)malloc(E->n_scalar * sizeof(float));
    E->warped = (float *)malloc(E->n_scalar * sizeof(float));
    E->gradient = (field_t *)calloc(E->n_vector, sizeof(field_t));
    E->update = (field_t *)calloc(E->n_vector, sizeof(field_t));
    E->deformation = (field_t *)calloc(E->n_vector, sizeof(field_t)); /* D starts at 0, cc:186 */
    E->mse_term = (double *)calloc(E->n_vox, sizeof(double));
    E->valid = (int64_t *)calloc(E->n_vox, sizeof(int64_t));
    memcpy(E->fixed, fixed, E->n_scalar * sizeof(float));
    memcpy(E->moving, moving, E->n_scalar * sizeof(float));
    memcpy(E->warped, moving, E->n_scalar * sizeof(float)); /* W0 := M, not a zero-field warp, cc:183 */
    orc_gradient(E);                                         /* once, of the fixed image, cc:74 */
    return E;
}

void ORC_NAME(orc_engine_destroy)(void *e) {
    orc_engine *E = (orc_engine *)e;
    if (!E) return;
    free(E->fixed);
    free(E->moving);
    free(E->warped);
    free(E->gradient);
    free(E->update);
    free(E->deformation);
    free(E->mse_term);
    free(E->valid);
    free(E);
}

void ORC_NAME(orc_engine_compute_update_and_mse)(void *e, double *mse, int64_t *n) {
    orc_update_and_mse((orc_engine *)e, mse, n);
}
void ORC_NAME(orc_engine_smooth_update)(void *e, double s) { orc_smooth((orc_engine *)e, ((orc_engine *)e)->update, s); }
void ORC_NAME(orc_engine_smooth_deformation)(void *e, double s) {
    orc_smooth((orc_engine *)e, ((orc_engine *)e)->deformation, s);
}
void ORC_NAME(orc_engine_add_update)(void *e) { orc_add_update((orc_engine *)e); }
void ORC_NAME(orc_engine_add_update_diffeomorphic)(void *e) { orc_add_update_diffeomorphic((orc_engine *)e); }
void ORC_NAME(orc_engine_warp_moving)(void *e) { orc_warp_moving((orc_engine *)e); }

/* compute_single_iteration, src/Alignment_Demons.cc:81-108 */
double ORC_NAME(orc_engine_iterate)(void *e) {
    orc_engine *E = (orc_engine *)e;
    double mse = 0.0;
    int64_t n = 0;
    orc_update_and_mse(E, &mse, &n);
    if (E->p.use_diffeomorphic && E->p.update_field_smoothing_sigma > 0.0)
        orc_smooth(E, E->update, E->p.update_field_smoothing_sigma);
    if (!E->p.use_diffeomorphic)
        orc_add_update(E);
    else
        orc_add_update_diffeomorphic(E);
    if (E->p.deformation_field_smoothing_sigma > 0.0) orc_smooth(E, E->deformation, E->p.deformation_field_smoothing_sigma);
    orc_warp_moving(E);
    return mse;
}

/* which: 0 fixed, 1 moving, 2 warped (float[N*C]); 3 gradient, 4 update, 5 deformation (as double[N*3]) */
int ORC_NAME(orc_engine_get)(void *e, int which, void *out) {
    orc_engine *E = (orc_engine *)e;
    const float *s = NULL;
    const field_t *v = NULL;
    switch (which) {
        case 0: s = E->fixed; break;
        case 1: s = E->moving; break;
        case 2: s = E->warped; break;
        case 3: v = E->gradient; break;
        case 4: v = E->update; break;
        case 5: v = E->deformation; break;
        default: return 1;
    }
    if (s) memcpy(out, s, E->n_scalar * sizeof(float));
    if (v)
        for (size_t i = 0; i < E->n_vector; ++i) ((double *)out)[i] = (double)v[i];
    return 0;
}
int ORC_NAME(orc_engine_set)(void *e, int which, const void *in) {
    orc_engine *E = (orc_engine *)e;
    float *s = NULL;
    field_t *v = NULL;
    switch (which) {
        case 0: s = E->fixed; break;
        case 1: s = E->moving; break;
        case 2: s = E->warped; break;
        case 3: v = E->gradient; break;
        case 4: v = E->update; break;
        case 5: v = E->deformation; break;
        default: return 1;
    }
    if (s) memcpy(s, in, E->n_scalar * sizeof(float));
    if (v)
        for (size_t i = 0; i < E->n_vector; ++i) v[i] = (field_t)((const double *)in)[i];
    return 0;
}

/* Driver loop of AlignViaDemons, src/Alignment_Demons.cc:977-997 (on raw volumes). */
int ORC_NAME(orc_demons_run)(const orc_geom *g, const float *fixed, const float *moving, const orc_params *p,
                             double *deformation_out, double *mse_history, int64_t *iterations_done) {
    orc_engine *E = (orc_engine *)ORC_NAME(orc_engine_create)(g, fixed, moving, p);
    if (!E) return 1;
    double prev_mse = INFINITY; /* cc:980 */
    int64_t done = 0;
    for (int64_t iter = 0; iter < p->max_iterations; ++iter) {
        const double mse = ORC_NAME(orc_engine_iterate)(E);
        if (mse_history) mse_history[iter] = mse;
        done = iter + 1;
        const double mse_change = fabs(prev_mse - mse);
        if (mse_change < p->convergence_threshold && iter > 0) break; /* cc:987-993 */
        prev_mse = mse;
    }
    if (iterations_done) *iterations_done = done;
    ORC_NAME(orc_engine_get)(E, 5, deformation_out); /* cc:997 */
    ORC_NAME(orc_engine_destroy)(E);
    return 0;
}
