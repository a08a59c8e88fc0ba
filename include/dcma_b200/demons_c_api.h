/* dcma_b200 -- C ABI of the B200-native Demons registration path (drop-in for DICOMautomaton's Alignment_Demons).
 *
 * The reference has no FFI for this path: `AlignViaDemons` (src/Alignment_Demons.h:233-236) constructs a file-local
 * SYCL engine (src/Alignment_Demons.cc:52-704) in-process.  This header is the boundary a maintainer binds instead;
 * every entry point names the reference interface it replaces.  Plain C, plain pointers and sizes, no C++/torch types,
 * nothing throws across it: every call returns 0 on success or a negative dcma_b200_status and leaves a message in
 * the caller's error buffer (and in dcma_b200_last_error()).
 *
 * There is NO CPU fallback: without a CUDA device (sm_100a) every compute entry point fails with DCMA_B200_ENODEVICE.
 *
 * Layouts (identical to the reference): images are dense row-major  idx = ((z*rows + y)*cols + x)*channels + c
 * (src/Alignment_Demons.h:77-79), float; vector fields handed across this boundary are double, 3 interleaved
 * components per voxel in mm along the index axes (x=column, y=row, z=slice), pull convention
 * warped(p) = moving(p + D(p)) (src/Alignment_Demons.h:217-220).
 */
#ifndef DCMA_B200_DEMONS_C_API_H
#define DCMA_B200_DEMONS_C_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCMA_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define DCMA_B200_API __attribute__((visibility("default")))
#else
#define DCMA_B200_API
#endif

typedef enum {
    DCMA_B200_OK = 0,
    DCMA_B200_EINVAL = -1,    /* bad argument (reference: std::invalid_argument) */
    DCMA_B200_ENODEVICE = -2, /* no usable CUDA device / wrong architecture */
    DCMA_B200_ECUDA = -3,     /* CUDA runtime error (reference: SYCL async exception, cc:140-150) */
    DCMA_B200_ENOMEM = -4,
    DCMA_B200_ESTATE = -5     /* call made in the wrong engine state */
} dcma_b200_status;

/* Header of `demons_volume<T>` (src/Alignment_Demons.h:64-74). */
typedef struct {
    int64_t slices, rows, cols, channels;
    double pxl_dx, pxl_dy, pxl_dz; /* mm: x=column spacing, y=row spacing, z=slice spacing */
} dcma_b200_geom;

/* How the gradient/update/deformation fields are held in HBM. */
typedef enum {
    DCMA_B200_FIELD_FP64 = 0,        /* double storage + double arithmetic with FMA contraction (default) */
    DCMA_B200_FIELD_FP64_STRICT = 1, /* double storage, no contraction, IEEE division: bit-identical deformation field
                                        to the reference's x86-64 build */
    DCMA_B200_FIELD_FP32 = 2,        /* float storage, fp32 smoothing; coordinates/interpolation/force stay fp64 */
    DCMA_B200_FIELD_MIXED = 3        /* deformation field and all arithmetic as DCMA_B200_FIELD_FP64; only the per-iteration
                                        update field is stored (and smoothed) as float: 176 instead of 256 algorithmic
                                        bytes per voxel-update, within 1e-6 mm of the reference after 200 iterations */
} dcma_b200_field_mode;

/* POD mirror of `AlignViaDemonsParams` (src/Alignment_Demons.h:20-61), same names, same defaults
 * (dcma_b200_demons_params_default), followed by additive extension fields whose defaults keep reference behaviour. */
typedef struct {
    int64_t max_iterations;                   /* 100 */
    double convergence_threshold;             /* 0.001 */
    double deformation_field_smoothing_sigma; /* 1.0 mm */
    double update_field_smoothing_sigma;      /* 0.5 mm; only used when use_diffeomorphic (cc:86) */
    int32_t use_diffeomorphic;                /* 0 */
    int32_t use_histogram_matching;           /* 0; host-side pre-step, ignored by this library (cc:967-972) */
    int64_t histogram_bins;                   /* 256; ditto */
    double histogram_outlier_fraction;        /* 0.01; ditto */
    double normalization_factor;              /* 1.0 */
    double max_update_magnitude;              /* 2.0 */
    int64_t verbosity;                        /* 1; >=1 prints one line per iteration to stderr (cc:983-985) */
    /* ---- extensions (absent in the reference) ---- */
    int32_t field_mode;          /* dcma_b200_field_mode; default DCMA_B200_FIELD_FP64 */
    int32_t device;              /* CUDA device ordinal; -1 = current device */
    int32_t check_every;         /* iterations between host polls of the device-side convergence flag; 0 = auto */
    int32_t pyramid_levels;      /* 1 = reference behaviour (single resolution); 2..4: dcma_b200_demons_register runs a
                                    coarse-to-fine pyramid (definition: dicomautomaton_b200/csrc/pyramid.cu) */
    int32_t use_symmetric_force; /* 0 = reference behaviour (fixed-image gradient only); 1: the force uses
                                    J = (grad F + grad W) / 2, grad W by the rule of cc:232-265 applied to the warped image */
    int32_t fused;               /* 1 (default) = fused loop kernels; 0 = one kernel per reference stage */
    int32_t n_devices;           /* dcma_b200_demons_register only: > 1 shards the volume into z-slabs over devices
                                    0..n_devices-1 of this process (0/1 = single device, reference-like) */
    int32_t halo_planes;         /* z-slab halo; 0 = default */
    int32_t pyramid_iterations[3]; /* iterations at pyramid levels 1..3 (coarser than the finest); 0 = max_iterations */
    int32_t reserved[1];
} dcma_b200_demons_params;

/* Timing/launch statistics of one run (all device-timed with CUDA events on the engine's stream). */
typedef struct {
    int64_t iterations_done; /* number of compute_single_iteration equivalents executed (cc:981-995) */
    int64_t kernel_launches; /* kernels launched by this library inside the loop */
    double loop_ms;          /* iteration loop only */
    double h2d_ms, d2h_ms;   /* host<->device copies of the one-shot call (0 for engine-level runs) */
    double stage_ms[8];      /* per-stage totals when profiling is on: see dcma_b200_stage */
    int64_t stage_launches[8];
    int64_t h2d_bytes, d2h_bytes;
} dcma_b200_run_stats;

typedef enum {
    DCMA_B200_STAGE_FORCE = 0,     /* compute_update_and_mse            cc:274-343 (fused: warp+force+MSE) */
    DCMA_B200_STAGE_SMOOTH_U = 1,  /* smooth_on_device(update)          cc:563-702 */
    DCMA_B200_STAGE_ADD = 2,       /* add_update                        cc:345-369 */
    DCMA_B200_STAGE_COMPOSE = 3,   /* add_update_diffeomorphic          cc:371-465 */
    DCMA_B200_STAGE_SMOOTH_D = 4,  /* smooth_on_device(deformation)     cc:563-702 */
    DCMA_B200_STAGE_WARP = 5,      /* warp_moving                       cc:467-550 */
    DCMA_B200_STAGE_REDUCE = 6,    /* host MSE loop + convergence test  cc:334-342, 987-993 */
    DCMA_B200_STAGE_GRADIENT = 7   /* compute_gradient                  cc:208-272 (stage API only) */
} dcma_b200_stage;

typedef enum { /* buffers addressable through engine_get/engine_set (same numbering as the oracle harness) */
    DCMA_B200_BUF_FIXED = 0,
    DCMA_B200_BUF_MOVING = 1,
    DCMA_B200_BUF_WARPED = 2,   /* float[N*channels] */
    DCMA_B200_BUF_GRADIENT = 3, /* double[N*3], computed on demand */
    DCMA_B200_BUF_UPDATE = 4,   /* double[N*3] */
    DCMA_B200_BUF_DEFORMATION = 5
} dcma_b200_buffer;

typedef struct dcma_b200_engine dcma_b200_engine; /* opaque; replaces `sycl_demons_engine` (cc:52-704) */

/* ---- library ----------------------------------------------------------------------------------------------- */
DCMA_B200_API int dcma_b200_abi_version(void);
DCMA_B200_API const char *dcma_b200_last_error(void); /* thread-local message of the last failing call */
DCMA_B200_API int dcma_b200_device_count(void);       /* number of CUDA devices visible, 0 if none (never fails) */
DCMA_B200_API void dcma_b200_demons_params_default(dcma_b200_demons_params *p); /* defaults of Alignment_Demons.h:20-61 */
/* Engines return their device blocks to a process-lifetime pool instead of the driver, so that the next registration of
 * the same shape does not pay cudaMalloc / cudaFree of several GB again (10-25 ms per call, hundreds on a busy box).  The
 * pool keeps at most DCMA_B200_DEVICE_POOL_MB per device (environment, default 12288; 0 = no caching).  _trim gives
 * everything cached back to the driver; _bytes reports how much is cached.  The reference has no counterpart. */
DCMA_B200_API void dcma_b200_device_pool_trim(void);
DCMA_B200_API int64_t dcma_b200_device_pool_bytes(void);

/* ---- one-shot registration on HOST buffers ----------------------------------------------------------------- */
/* Replaces, inside AlignViaDemons, the block  `sycl_demons_engine engine(fixed_vol, moving_vol, params); for(iter...)
 * { engine.compute_single_iteration(); ... } engine.export_deformation_volume()`  (src/Alignment_Demons.cc:977-997).
 * `moving_on_fixed_grid` is the moving image after resample_image_to_reference_grid / histogram_match (cc:964-972).
 * deformation_out: N*3 doubles (AoS z,y,x,c).  mse_history: max_iterations doubles (with pyramid_levels > 1: the sum of
 * the levels' iteration limits, coarse levels first) or NULL.  stats may be NULL.
 * Blocking; H2D and D2H copies happen inside the call. */
DCMA_B200_API int dcma_b200_demons_register(const dcma_b200_geom *geom, const float *fixed, const float *moving_on_fixed_grid,
                              const dcma_b200_demons_params *params, double *deformation_out, double *mse_history,
                              dcma_b200_run_stats *stats, char *errbuf, size_t errbuf_len);

/* ---- engine API (device-resident; used by the benchmark, the parity tests and the slab/batch drivers) ------- */
/* Constructor of sycl_demons_engine (cc:54-75): allocates device buffers for `geom`. `stream` is a cudaStream_t
 * (NULL = the library creates its own non-blocking stream). */
DCMA_B200_API int dcma_b200_engine_create(dcma_b200_engine **out, const dcma_b200_geom *geom, const dcma_b200_demons_params *params,
                            void *stream, char *errbuf, size_t errbuf_len);
DCMA_B200_API void dcma_b200_engine_destroy(dcma_b200_engine *e); /* cc:77-79 */

/* allocate_device_memory's copies (cc:181-186): F, M uploaded, W := M, D := 0, U := 0, iteration counter := 0.
 * `on_device` != 0 means the pointers are device pointers on the engine's device (no PCIe traffic). */
DCMA_B200_API int dcma_b200_engine_load(dcma_b200_engine *e, const float *fixed, const float *moving_on_fixed_grid, int on_device);

/* The loop of AlignViaDemons (cc:980-995) for at most `max_iterations` further iterations (continues from the
 * current state; convergence is evaluated on the device with the reference's rule). `profile` != 0 brackets every
 * kernel with CUDA events and fills stats->stage_ms. mse_history: HOST array of max_iterations doubles or NULL. */
DCMA_B200_API int dcma_b200_engine_run(dcma_b200_engine *e, int64_t max_iterations, int profile, double *mse_history,
                         dcma_b200_run_stats *stats);

/* One reference engine method at a time (stage-level parity; same order of buffers as the reference):
 * FORCE reads `warped` and writes `update` and returns (mse, n_valid) through the two pointers. */
DCMA_B200_API int dcma_b200_engine_stage(dcma_b200_engine *e, int stage, double sigma_mm, double *mse_out, int64_t *n_valid_out);

/* export_deformation_volume (cc:110-113) and friends; `host` is a host pointer sized as in dcma_b200_buffer. */
DCMA_B200_API int dcma_b200_engine_get(dcma_b200_engine *e, int buffer, void *host);
DCMA_B200_API int dcma_b200_engine_set(dcma_b200_engine *e, int buffer, const void *host);
/* Same as get(DEFORMATION) but into a DEVICE pointer (double[N*3], AoS) -- for device-resident consumers. */
DCMA_B200_API int dcma_b200_engine_export_deformation_device(dcma_b200_engine *e, double *device_out);

/* Synchronise the engine's stream (cudaStreamSynchronize). */
DCMA_B200_API int dcma_b200_engine_sync(dcma_b200_engine *e);
/* Bytes of HBM the engine holds. */
DCMA_B200_API int64_t dcma_b200_engine_device_bytes(const dcma_b200_engine *e);

/* ---- one registration sharded into z-slabs over several GPUs (no counterpart in the reference, which keeps the whole
 * volume in one address space, src/Alignment_Demons.cc:171-187; SURVEY.md 8(e), BASELINE config 4) ------------------
 * Every slab engine holds its slices plus `halo_planes` planes per interior side; the moving image is replicated.  Per
 * iteration the z-neighbours exchange halo planes before each stencil/gather stage and all slabs gather their MSE terms;
 * the result equals the whole-volume engine's bit for bit.  halo_planes < 0 selects the default
 * (max(smoothing radius along z, 8)); a composition gather that leaves the halo fails the run with DCMA_B200_ESTATE. */
typedef struct dcma_b200_slab_group dcma_b200_slab_group;

/* Pure host helper: the slab of `rank` -- local slice 0 is global slice *zoff, the engine holds *local slices and owns
 * the local range [*own0, *own1). */
DCMA_B200_API int dcma_b200_slab_partition(int64_t slices, int world, int rank, int64_t halo_planes, int64_t *zoff,
                                           int64_t *local, int64_t *own0, int64_t *own1);
/* One process drives `n_devices` engines (device ordinals may repeat); halos move by peer copies over NVLink. */
DCMA_B200_API int dcma_b200_slab_group_create_local(dcma_b200_slab_group **out, const dcma_b200_geom *geom,
                                                    const dcma_b200_demons_params *params, const int32_t *devices,
                                                    int32_t n_devices, int64_t halo_planes, char *errbuf, size_t errbuf_len);
/* One rank per process (the torchrun layout): rank 0 calls dcma_b200_nccl_unique_id and broadcasts the 128 bytes by
 * any means; every rank then creates its part of the group.  Halos move by ncclSend/ncclRecv, MSE terms by
 * ncclAllGather.  NCCL is loaded at run time (libnccl.so.2, or the path in DCMA_B200_NCCL_LIB). */
DCMA_B200_API int dcma_b200_nccl_unique_id(unsigned char id[128], char *errbuf, size_t errbuf_len);
DCMA_B200_API int dcma_b200_slab_group_create_nccl(dcma_b200_slab_group **out, const dcma_b200_geom *geom,
                                                   const dcma_b200_demons_params *params, const unsigned char id[128],
                                                   int32_t rank, int32_t world, int32_t device, int64_t halo_planes,
                                                   char *errbuf, size_t errbuf_len);
DCMA_B200_API void dcma_b200_slab_group_destroy(dcma_b200_slab_group *g);
/* fixed / moving: WHOLE volumes (host pointers, or device pointers readable from every engine's device). */
DCMA_B200_API int dcma_b200_slab_group_load(dcma_b200_slab_group *g, const float *fixed, const float *moving_on_fixed_grid,
                                            int on_device);
/* The loop of AlignViaDemons (cc:980-995) over all slabs; mse_history/stats as dcma_b200_engine_run. */
DCMA_B200_API int dcma_b200_slab_group_run(dcma_b200_slab_group *g, int64_t max_iterations, double *mse_history,
                                           dcma_b200_run_stats *stats);
/* Slices [*z0, *z1) held by THIS process (everything for a local group) and their deformation, AoS doubles. */
DCMA_B200_API int dcma_b200_slab_group_range(dcma_b200_slab_group *g, int64_t *z0, int64_t *z1);
DCMA_B200_API int dcma_b200_slab_group_get_deformation(dcma_b200_slab_group *g, double *deformation_out);
DCMA_B200_API int64_t dcma_b200_slab_group_halo_bytes(const dcma_b200_slab_group *g); /* bytes exchanged by the last run */
/* Exchange statistics of the last run of THIS process: device time spent inside halo exchanges (ms, all of them summed;
 * with `overlapped` != 0 they ran on a second stream under the interior kernels), their number, the halo width in planes
 * and the transport (0 in-process peer copies, 1 NCCL send/recv, 2 IPC pull + NCCL tokens, 3 IPC push + flag counters).
 * Any pointer may be NULL.  No reference counterpart (the reference has one address space, src/Alignment_Demons.cc:171-187). */
DCMA_B200_API int dcma_b200_slab_group_exchange_stats(const dcma_b200_slab_group *g, double *exchange_ms, int64_t *exchanges,
                                                      int32_t *overlapped, int32_t *transport, int64_t *halo_planes);

/* ---- pyramid building blocks (extension; host buffers) ------------------------------------------------------- */
/* Next-coarser level of `image`: every axis with extent >= 8 is halved (mean of the finite voxels of each block).
 * *coarse receives the coarse geometry; `out` may be NULL to query the geometry only. */
DCMA_B200_API int dcma_b200_restrict_image(const dcma_b200_geom *fine, const float *image, dcma_b200_geom *coarse, float *out,
                                           int device, char *errbuf, size_t errbuf_len);
/* Trilinear prolongation of a displacement field (AoS doubles, mm) from the next-coarser level onto `fine`. */
DCMA_B200_API int dcma_b200_prolong_field(const dcma_b200_geom *fine, const double *coarse_field, double *out, int device,
                                          char *errbuf, size_t errbuf_len);

/* ---- warp an image with a deformation field on the same grid (warp_moving semantics, cc:467-550) ------------ */
/* out[p] = image(p + D(p)) with the reference's strict validity rule; `oob_value` replaces NaN when finite
 * (WarpImages uses 0.0, src/Operations/WarpImages.cc:197,452).  All pointers HOST. */
DCMA_B200_API int dcma_b200_warp_image(const dcma_b200_geom *geom, const float *image, const double *deformation, float oob_value,
                         int use_oob_value, float *out, int device, char *errbuf, size_t errbuf_len);

/* ---- pinned host memory ------------------------------------------------------------------------------------------ */
/* Page-locked host buffers for the volumes handed to the calls above: transfers from/to them run at PCIe speed and
 * asynchronously (H2D of the images, D2H of the 24 B/voxel field).  The reference has no counterpart: its engine copies
 * between std::vector and USM (src/Alignment_Demons.cc:116-121, 181-186).  Page-locking costs about as much as one
 * transfer saves, so callers keep such buffers for the life of the process (host/Alignment_Demons.cc pools them). */
DCMA_B200_API int dcma_b200_host_alloc(void **out, size_t bytes, char *errbuf, size_t errbuf_len);
DCMA_B200_API void dcma_b200_host_free(void *p);

/* ---- histogram matching (AlignViaDemonsHelpers::histogram_match, src/Alignment_Demons.cc:760-891) ------------- */
/* out := source with its intensity distribution matched to `reference` (num_bins-bin CDFs between the
 * outlier_fraction / 1-outlier_fraction percentiles, first-bin->=quantile lookup, clamping outside the source range).
 * Same double arithmetic as the reference: bit-identical output.  *mapped (may be NULL) is set to 0 when the source is
 * returned unchanged (no finite values, or a degenerate intensity range: cc:791-794, 812-815).  All pointers HOST;
 * the two images need not have the same size. */
DCMA_B200_API int dcma_b200_histogram_match(const float *source, int64_t n_source, const float *reference, int64_t n_reference,
                                            int64_t num_bins, double outlier_fraction, float *out, int32_t *mapped, int device,
                                            char *errbuf, size_t errbuf_len);

/* ---- warp an image that lives on ANOTHER regular grid than the field (WarpImages with a deformation_field) ----- */
/* A regular grid in world space: voxel (z, y, x) sits at origin + x*pxl_dx*ex + y*pxl_dy*ey + z*pxl_dz*ez, with ex, ey,
 * ez unit vectors (the reference's row_unit, col_unit and ortho_unit of a planar_image stack; origin = position(0,0) of
 * the first image). */
typedef struct {
    dcma_b200_geom geom;
    double origin[3];
    double ex[3], ey[3], ez[3];
} dcma_b200_grid;
/* ---- one-shot registration of two image STACKS, each on its own regular grid --------------------------------------- */
/* What AlignViaDemons does between marshalling and the construction of the deformation_field
 * (src/Alignment_Demons.cc:964-997) in ONE call: the moving stack is resampled onto the fixed stack's grid on the device
 * (resample_image_to_reference_grid, cc:709-755: trilinear, NaN outside the moving grid) and never travels back to the
 * host; then the loop; then the field comes back in chunks.  `on_field_chunk` (may be NULL) is called on the calling
 * thread each time voxels [voxel_begin, voxel_end) of deformation_out (N*3 doubles, AoS z,y,x,c) have landed, in
 * ascending order, while later chunks are still in flight -- the caller can copy them into its own image stack in the
 * meantime.  With pyramid_levels > 1 or n_devices > 1 the call is equivalent to dcma_b200_resample_to_grid +
 * dcma_b200_demons_register and the callback fires once at the end.  Histogram matching is the caller's pre-step and
 * needs the resampled image on the host: use the separate calls then.  All pointers HOST (pinned = asynchronous). */
typedef void (*dcma_b200_chunk_fn)(void *user, int64_t voxel_begin, int64_t voxel_end);
DCMA_B200_API int dcma_b200_demons_register_stacks(const dcma_b200_grid *fixed_grid, const float *fixed,
                                                   const dcma_b200_grid *moving_grid, const float *moving,
                                                   const dcma_b200_demons_params *params, double *deformation_out,
                                                   double *mse_history, dcma_b200_run_stats *stats,
                                                   dcma_b200_chunk_fn on_field_chunk, void *user, char *errbuf,
                                                   size_t errbuf_len);

/* The per-voxel body of WarpImages for a deformation_field transform (src/Operations/WarpImages.cc:418-456,
 * src/Alignment_Field.cc:138-153): out(p) = image(p + D(p)) with p running over `image_grid`, D sampled trilinearly
 * from the field on `field_grid` (displacement components along the field's axes, mm), the image sampled trilinearly at
 * the displaced position; `oob_value` (WarpImages: 0.0, cc:197) wherever either lookup leaves its grid or is non-finite.
 * `mask` (image_grid voxels, may be NULL): 0 = voxel outside the region of interest, copied unchanged.  All HOST. */
DCMA_B200_API int dcma_b200_warp_image_grid(const dcma_b200_grid *field_grid, const double *deformation,
                                            const dcma_b200_grid *image_grid, const float *image, const unsigned char *mask,
                                            float oob_value, float *out, int device, char *errbuf, size_t errbuf_len);

/* The same body with the operation's THREE stacks (src/Operations/WarpImages.cc:385-456): every voxel p of the OUTPUT stack
 * (the operation's copy of the ReferenceImageSelection, `out_inout` holds its voxels on entry) that `mask` selects (NULL =
 * all; the operation's bounded voxels, i.e. inside the selected ROIs) takes source(p + D(p)), the SOURCE stack (the
 * ImageSelection) sampled trilinearly; `oob_value` where a lookup leaves its grid.  channel >= 0: only that channel is
 * assigned (:422); < 0: all.  Unselected voxels and channels keep their values.  All pointers HOST. */
DCMA_B200_API int dcma_b200_warp_to_grid(const dcma_b200_grid *field_grid, const double *deformation,
                                         const dcma_b200_grid *source_grid, const float *source, const dcma_b200_grid *out_grid,
                                         const unsigned char *mask, int32_t channel, float oob_value, float *out_inout,
                                         int device, char *errbuf, size_t errbuf_len);

/* ---- synthetic deformation fields for testing a registration (the reference's generator operations) -------------- */
/* field_out[z][y][x][3] (mm, AoS doubles, `grid` voxels) := the field of GenerateVirtualDataSinusoidalFieldV1
 * (src/Operations/GenerateVirtualDataSinusoidalFieldV1.cc:77-193: three plane waves with 2, 3 and 1.5 periods across
 * the bounds, magnitudes rescaled to [0, 1]) or of GenerateVirtualDataVortexFieldV1
 * (src/Operations/GenerateVirtualDataVortexFieldV1.cc:75-178: a cylindrical vortex about the centre of the bounds,
 * divided by its largest magnitude).  bounds = {x_min, x_max, y_min, y_max, z_min, z_max}; the operations use
 * {0, 512, 0, 512, 0, 100} on a 512 x 512 x 100 grid whose slice z sits at z_min + pxl_dz * (z + 1/2).  extremes (may be
 * NULL) := {min, max} displacement magnitude before normalisation (the operations record them as metadata).  HOST. */
typedef enum { DCMA_B200_VFIELD_SINUSOIDAL_V1 = 0, DCMA_B200_VFIELD_VORTEX_V1 = 1 } dcma_b200_vfield_kind;
DCMA_B200_API int dcma_b200_virtual_field(int32_t kind, const dcma_b200_grid *grid, const double *bounds, double *field_out,
                                          double *extremes, int device, char *errbuf, size_t errbuf_len);

/* resample_image_to_reference_grid (src/Alignment_Demons.cc:709-755): the source image sampled trilinearly at the
 * positions of the destination grid's voxels, `oob_value` (the reference passes NaN, cc:743) outside the source grid.
 * out: dst voxels x src channels.  The reference's lookup is Ygor's (unpinned); definition as dcma_b200_warp_image_grid. */
DCMA_B200_API int dcma_b200_resample_to_grid(const dcma_b200_grid *src_grid, const float *image, const dcma_b200_grid *dst_grid,
                                             float oob_value, float *out, int device, char *errbuf, size_t errbuf_len);

/* ---- the numeric payload of a stored deformation field (the `.trans` text format) --------------------------------- */
/* deformation_field::write_to (src/Alignment_Field.cc:322-357) emits, after each image's header lines, every double of
 * the image on its own line at precision max_digits10 (`os << val << '\n'`, :346-348); read_from (:359-436) extracts
 * them again with `is >> val` (:418-420).  For a 512 x 512 x 256 field that is 2.0e8 numbers, ~4.8 GB of text.  These
 * two calls do that part on the GPU, byte-for-byte and bit-for-bit what the C library produces (correctly rounded in
 * both directions); the header lines stay with the caller (host/Field_Payload.cc is the stream-level codec that
 * patches/Alignment_Field.patch installs into the reference's two loops).
 *
 * dcma_b200_format_doubles: text := printf("%.17g\n") of values[0..n).  `capacity` bytes are available at `text`
 *   (25 * n always suffices; DCMA_B200_EINVAL if it turns out too small); *text_len := bytes written.  No terminator.
 * dcma_b200_parse_doubles:  values[0..n) := the first n white-space separated decimal numbers of text[0..text_len), as
 *   n stream extractions would read them; *consumed := offset just past the last character of the n-th number.
 *   DCMA_B200_EINVAL (message as the reference's "Pixel data could not be read") when the text holds fewer than n
 *   numbers, or a token is not a decimal number ("inf"/"nan" included, which the stream extraction rejects too) or
 *   overflows a double.  All pointers HOST. */
DCMA_B200_API int dcma_b200_format_doubles(const double *values, int64_t n, char *text, int64_t capacity, int64_t *text_len,
                                           int device, char *errbuf, size_t errbuf_len);
DCMA_B200_API int dcma_b200_parse_doubles(const char *text, int64_t text_len, int64_t n, double *values, int64_t *consumed,
                                          int device, char *errbuf, size_t errbuf_len);

#ifdef __cplusplus
}
#endif
#endif /* DCMA_B200_DEMONS_C_API_H */
