// Warp_Images_Field.h -- GPU form of the per-voxel body of the WarpImages operation for a deformation_field transform.
//
// Reference (src/Operations/WarpImages.cc:385-468): the operation copies the reference image array as a geometry
// placeholder, and PartitionedImageVoxelVisitorMutator calls, for every voxel bounded by the selected ROIs,
//     ref_p  = img.position(row, col)                                      (:424)
//     corr_p = ref_p; t.apply_to(corr_p)   // = ref_p + D(ref_p), three trilinear field look-ups   (:442-443, Alignment_Field.cc:138-153)
//     voxel  = img_adj.trilinearly_interpolate(corr_p, chan, 0.0)           (:452, InaccessibleValue :197)
// i.e. two adjacency-index walks and four trilinear interpolations per voxel on one host thread per image.  Here the ROI
// rasterisation stays the operation's own (its mutator marks the bounded voxels, patches/WarpImages.patch); everything
// per voxel after that is ONE call: dcma_b200_warp_to_grid (include/dcma_b200/demons_c_api.h).
#pragma once
#include <cstdint>
#include <vector>

#include "YgorImages.h"

#include "Alignment_Field.h"

namespace dcma_b200_host {

// field    : the transform (already a pull transform: WarpImages.cc:348-351)
// source   : the ImageSelection's stack (what img_adj indexes, :315)
// target   : the operation's copy of the reference image array (:385).  Bounded voxels of the selected channel(s) are
//            overwritten; everything else keeps the placeholder's values, as with the reference's functor.
// bounded  : one flag per voxel of `target` -- images in list order, row-major within an image; non-zero = bounded by
//            the ROIs.  Empty = every voxel.
// channel  : < 0 = all channels (:422)
// Returns false WITHOUT touching `target` when the three stacks cannot be described as regular grids in list order
// (the caller then keeps the reference's per-voxel path).  Throws std::runtime_error when the GPU call fails: a request
// the GPU path has accepted has no CPU fallback.
bool warp_stack_with_field(const deformation_field &field, const planar_image_collection<float, double> &source,
                           planar_image_collection<float, double> &target, const std::vector<unsigned char> &bounded,
                           int64_t channel, float inaccessible_value, int device = -1);

}  // namespace dcma_b200_host
