#!/usr/bin/env python3
"""Write patches/WarpImages.patch: the reference's src/Operations/WarpImages.cc with the per-voxel work of a
deformation-field warp handed to the GPU (dicomautomaton_b200/host/Warp_Images_Field.h), as a unified diff.  Needs
/root/reference; the patch itself is what is committed.

    python tools/make_warp_patch.py            (re-run when the reference file changes)

Two hunks, anchored on text: one include, and a block in front of the operation's own Process_Images_Parallel call.  The
operation's mutator still decides WHICH voxels are warped (ROIs, inclusivity, contour overlap): a first pass with a
functor that only marks bounded voxels replaces its per-voxel interpolation, then one GPU call fills them in.  Affine
transforms, stacks that are not regular grids, and anything else the GPU path declines keep the reference's own path."""
import difflib
import os

REF = os.environ.get("DCMA_REFERENCE_SRC", "/root/reference/src")
SRC = os.path.join(REF, "Operations", "WarpImages.cc")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "patches", "WarpImages.patch")

INCLUDE_ANCHOR = '#include "WarpImages.h"\n'
INCLUDE = '#include "../Warp_Images_Field.h"  //B200: per-voxel body of a deformation-field warp on the GPU.\n'

CALL_ANCHOR = '''                if(!edit_ia_ptr->imagecoll.Process_Images_Parallel( GroupIndividualImages,
                                                                    PartitionedImageVoxelVisitorMutator,
                                                                    {}, cc_ROIs, &ud )){
                    throw std::runtime_error("Unable to warp image array");
                }
'''

CALL = '''                // B200: a deformation-field warp of regular stacks runs on the GPU.  The mutator still decides which voxels
                // the ROIs bound (same inclusivity and contour-overlap options): a first pass only MARKS them in a scratch
                // copy of the placeholder, then one call does the two trilinear look-ups per voxel.
                bool warped_on_gpu = false;
                if(std::holds_alternative<deformation_field>( t_inv.transform )){
                    auto mark_ia = *edit_ia_ptr;
                    for(auto &mimg : mark_ia.imagecoll.images) mimg.fill_pixels(0.0f);
                    auto ud_mark = ud;
                    ud_mark.f_bounded = [](int64_t, int64_t, int64_t,
                                           std::reference_wrapper<planar_image<float,double>>,
                                           std::reference_wrapper<planar_image<float,double>>,
                                           float &voxel_val) { voxel_val = 1.0f; };
                    if(!mark_ia.imagecoll.Process_Images_Parallel( GroupIndividualImages,
                                                                   PartitionedImageVoxelVisitorMutator,
                                                                   {}, cc_ROIs, &ud_mark )){
                        throw std::runtime_error("Unable to mark the voxels to warp");
                    }
                    std::vector<unsigned char> bounded;
                    for(const auto &mimg : mark_ia.imagecoll.images){
                        for(int64_t r = 0; r < mimg.rows; ++r){
                            for(int64_t c = 0; c < mimg.columns; ++c){
                                bool any = false;
                                for(int64_t ch = 0; ch < mimg.channels; ++ch) any = any || (mimg.value(r, c, ch) != 0.0f);
                                bounded.push_back(any ? 1 : 0);
                            }
                        }
                    }
                    warped_on_gpu = dcma_b200_host::warp_stack_with_field( std::get<deformation_field>( t_inv.transform ),
                                                                           (*iap_it)->imagecoll, edit_ia_ptr->imagecoll,
                                                                           bounded, Channel, InaccessibleValue );
                }

                if(!warped_on_gpu
                && !edit_ia_ptr->imagecoll.Process_Images_Parallel( GroupIndividualImages,
                                                                    PartitionedImageVoxelVisitorMutator,
                                                                    {}, cc_ROIs, &ud )){
                    throw std::runtime_error("Unable to warp image array");
                }
'''


def patched(text: str) -> str:
    for a in (INCLUDE_ANCHOR, CALL_ANCHOR):
        if text.count(a) != 1:
            raise SystemExit(f"anchor not found exactly once: {a!r}")
    return text.replace(INCLUDE_ANCHOR, INCLUDE_ANCHOR + INCLUDE).replace(CALL_ANCHOR, CALL)


def main():
    if not os.path.isfile(SRC):
        raise SystemExit(f"{SRC} not found")
    old = open(SRC).read()
    diff = difflib.unified_diff(old.splitlines(True), patched(old).splitlines(True), "a/src/Operations/WarpImages.cc",
                                "b/src/Operations/WarpImages.cc", n=3)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.writelines(diff)
    print(OUT)


if __name__ == "__main__":
    main()
