// dcma_b200 -- precision-mode independent engine interface (host side).  The C ABI (c_api.cu) talks to this only.
#pragma once
#include <cstdint>

#include "../../include/dcma_b200/demons_c_api.h"

namespace dcma {

// Which slices of the volume an engine holds and computes (z-slab sharding).  The default is the whole
// volume.  Local slice l is global slice zoff + l; the engine holds `local` slices and computes [own0, own1) of them.
struct SlabSpec {
    int64_t zoff = 0, local = 0, own0 = 0, own1 = 0;  // local == 0: whole volume
};

// The stages of one loop iteration, for drivers that interleave halo exchanges (slab.cu).
enum class IterStage { Force, SmoothU, Update, SmoothD };

// Device-resident replacement of the reference's `sycl_demons_engine` (src/Alignment_Demons.cc:52-704).
class EngineBase {
  public:
    virtual ~EngineBase() {}
    // ---- z-slab drivers only ----
    virtual void begin_run(int64_t max_iterations) = 0;                       // reset the loop state, allocate the history
    virtual void iter_stage(IterStage st, long long it, bool global_mse) = 0;  // enqueue one stage of iteration `it`
    virtual void force_range(long long it, int z0, int z1) = 0;   // FORCE on owned slices [z0, z1) (local); no reduction
    virtual void force_finish(long long it) = 0;                  // reduces the MSE partials of the force_range calls
    virtual void update_range(long long it, int z0, int z1) = 0;  // diffeomorphic composition on owned slices [z0, z1)
    virtual bool symmetric() const = 0;
    virtual void publish_mse(double *dev_out2) = 0;                           // (sum, count) of the last FORCE -> 2 doubles
    virtual void stop_rule(const double *dev_gather, int world, long long it) = 0;
    virtual void end_run(double *mse_history_host, dcma_b200_run_stats *stats) = 0;
    virtual bool converged_poll() = 0;      // synchronising read of the device-side convergence flag
    virtual bool halo_overflowed() = 0;     // synchronising read
    virtual bool update_smoothed() const = 0, deformation_smoothed() const = 0, diffeomorphic() const = 0;
    virtual int smoothing_radius_z(bool update_field) const = 0;
    virtual void *field_base(int buffer) = 0;     // DCMA_B200_BUF_UPDATE / _DEFORMATION: device pointer of component 0
    virtual int field_allocation_count() const = 0;  // 3, or 4 when U and D differ in type (DCMA_B200_FIELD_MIXED)
    virtual void *field_allocation(int i) = 0;    // the field allocations (fixed identity; their roles rotate)
    virtual int64_t field_elem_size(int buffer) const = 0;  // 4 or 8, of DCMA_B200_BUF_UPDATE / _DEFORMATION
    virtual int64_t field_comp_stride() const = 0, plane_elems() const = 0;
    virtual void *stream_handle() = 0;
    virtual void get_owned_deformation(double *host_aos) = 0;  // owned slices only, AoS doubles
    virtual SlabSpec slab() const = 0;
    virtual void load(const float *fixed, const float *moving, bool on_device) = 0;
    // HOST pointers to whole volumes; of the moving image only slices [mz0, mz1) are uploaded -- the caller fills the
    // rest of moving_dev_mutable() itself (slab groups: from the other ranks over NVLink) before the first run
    virtual void load_partial_moving(const float *fixed, const float *moving, int64_t mz0, int64_t mz1) = 0;
    virtual float *moving_dev_mutable() = 0;
    virtual void run(int64_t max_iterations, bool profile, double *mse_history_host, dcma_b200_run_stats *stats) = 0;
    virtual void stage(int stage, double sigma_mm, double *mse_out, int64_t *n_valid_out) = 0;
    virtual void get(int buffer, void *host) = 0;
    // the deformation field (AoS doubles, owned slices) with a notification per landed chunk: `on_chunk(user, v0, v1)` is
    // called on the calling thread when voxels [v0, v1) of `host` are complete, ascending, while later chunks are in flight
    virtual void get_deformation_chunked(double *host, void (*on_chunk)(void *, int64_t, int64_t), void *user) = 0;
    virtual void set(int buffer, const void *host) = 0;
    virtual void export_deformation_device(double *device_out) = 0;
    virtual void import_deformation_device(const double *device_aos) = 0;  // warm start: D := field, W := warp(M, D)
    // the same for a slab engine: `device_aos_whole` is the field of the WHOLE volume (any device this one can read);
    // every slice the engine holds is taken from it, halo planes included
    virtual void import_deformation_whole(const double *device_aos_whole) = 0;
    virtual const float *fixed_dev() const = 0;
    virtual const float *moving_dev() const = 0;
    virtual void sync() = 0;
    virtual int64_t device_bytes() const = 0;
    virtual int device() const = 0;
};

// one factory per precision-mode translation unit
// `g` is always the geometry of the WHOLE volume; `slab` selects the part this engine holds.
EngineBase *make_engine_f64(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream, const SlabSpec &slab);
EngineBase *make_engine_f64_strict(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream, const SlabSpec &slab);
EngineBase *make_engine_f32(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream, const SlabSpec &slab);
EngineBase *make_engine_mixed(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream, const SlabSpec &slab);
EngineBase *make_engine(const dcma_b200_geom &g, const dcma_b200_demons_params &p, void *stream, const SlabSpec &slab);

// warp_moving semantics on host buffers (strict fp64 translation unit)
void warp_image_host(const dcma_b200_geom &g, const float *image, const double *deformation_aos, bool use_oob,
                     float oob_value, float *out, int device);

// histogram_match (cc:760-891) on the GPU, bit-identical (histmatch.cu); returns 1 if mapped, 0 if returned unchanged
int histogram_match_host(const float *source, long long n_src, const float *reference, long long n_ref, long long num_bins,
                         double outlier_fraction, float *out, int device);

// WarpImages' voxel body for a field and an image on different regular grids (warp_grid.cu)
void warp_image_grid_host(const dcma_b200_grid &field_grid, const double *deformation_aos, const dcma_b200_grid &image_grid,
                          const float *image, const unsigned char *mask, float oob_value, float *out, int device);

// the same with separate source and output stacks, ROI mask and channel selection, as the operation has them
void warp_to_grid_host(const dcma_b200_grid &field_grid, const double *deformation_aos, const dcma_b200_grid &source_grid,
                       const float *source, const dcma_b200_grid &out_grid, const unsigned char *mask, int channel, float oob_value,
                       float *out_inout, int device);

// device-pointer form of the same kernel, on `stream` (void* = cudaStream_t); out: dst voxels x src channels
void resample_to_grid_device(const dcma_b200_grid &src_grid, const float *d_image, const dcma_b200_grid &dst_grid, float oob_value,
                             float *d_out, void *stream);
void resample_to_grid_host(const dcma_b200_grid &src_grid, const float *image, const dcma_b200_grid &dst_grid, float oob_value,
                           float *out, int device);

// the two synthetic deformation fields of the reference's generator operations (virtual_field.cu); extremes[2] (may be
// null) receives the minimum and maximum displacement magnitude before normalisation
void virtual_field_host(int kind, const dcma_b200_grid &grid, const double bounds[6], double *out_aos, double extremes[2], int device);

// the numeric payload of deformation_field::write_to / read_from (src/Alignment_Field.cc:322-436) on the GPU (text_io.cu)
long long format_doubles_host(const double *values, long long n, char *out, long long capacity, int device);
long long parse_doubles_host(const char *text, long long nbytes, long long n, double *out, int device);

}  // namespace dcma
