// Stand-in for src/YgorImages_Functors/Processing/Partitioned_Image_Voxel_Visitor_Mutator.h: the option and user-data types
// WarpImages.cc fills in, and a mutator that visits every voxel of one image and calls f_bounded / f_unbounded.
// "Bounded" here: the voxel CENTRE lies inside a closed contour of the selected collections whose plane is within half a
// slice thickness of the image plane (crossing-number test in the image plane); overlapping contours form a union.  The
// reference's mutator (Ygor's Mutate_Voxels) also implements the other inclusivity and overlap modes; this stand-in
// accepts the options and applies the centre / union rule for all of them.  NOT a product component.
#pragma once
#include <any>
#include <cmath>
#include <cstdint>
#include <functional>
#include <list>
#include <string>

#include "YgorImages.h"
#include "YgorMath.h"

template <class T, class R>
using Mutate_Voxels_Functor = std::function<void(int64_t, int64_t, int64_t, std::reference_wrapper<planar_image<T, R>>,
                                                 std::reference_wrapper<planar_image<T, R>>, T &)>;

struct Mutate_Voxels_Opts {
    enum class EditStyle { InPlace, Surrogate } editstyle = EditStyle::InPlace;
    enum class Aggregate { Mean, Median, First } aggregate = Aggregate::First;
    enum class Adjacency { SingleVoxel, NearestNeighbours } adjacency = Adjacency::SingleVoxel;
    enum class MaskMod { Union, Noop } maskmod = MaskMod::Noop;
    enum class ContourOverlap { Ignore, HonourOppositeOrientations, ImplicitOrientations } contouroverlap = ContourOverlap::Ignore;
    enum class Inclusivity { Centre, Inclusive, Exclusive } inclusivity = Inclusivity::Centre;
};

struct PartitionedImageVoxelVisitorMutatorUserData {
    Mutate_Voxels_Opts mutation_opts;
    std::string description;
    Mutate_Voxels_Functor<float, double> f_bounded, f_unbounded, f_visitor;
};

inline bool PartitionedImageVoxelVisitorMutator(std::list<planar_image<float, double>>::iterator first_img_it,
                                                std::list<std::list<planar_image<float, double>>::iterator> /*selected*/,
                                                std::list<std::reference_wrapper<planar_image_collection<float, double>>> /*external*/,
                                                std::list<std::reference_wrapper<contour_collection<double>>> ccs, std::any user_data) {
    auto *ud = std::any_cast<PartitionedImageVoxelVisitorMutatorUserData *>(user_data);
    if (!ud) return false;
    auto &img = *first_img_it;
    planar_image<float, double> mask_img = img;
    for (int64_t r = 0; r < img.rows; ++r)
        for (int64_t c = 0; c < img.columns; ++c) {
            const auto p = img.position(r, c);
            bool inside = false;
            for (const auto &cc : ccs)
                for (const auto &cop : cc.get().contours) {
                    if (inside || cop.points.size() < 3) continue;
                    double fr, fc, dist;
                    img.fractional(cop.points.front(), fr, fc, dist);
                    if (std::abs(dist) > 0.5 * img.pxl_dz) continue;
                    double pr, pc, pd;
                    img.fractional(p, pr, pc, pd);
                    bool in = false;  // crossing number in (col, row) coordinates
                    auto prev = cop.points.back();
                    for (const auto &cur : cop.points) {
                        double ar, ac, br, bc, d;
                        img.fractional(prev, ar, ac, d);
                        img.fractional(cur, br, bc, d);
                        if (((ar > pr) != (br > pr)) && (pc < (bc - ac) * (pr - ar) / (br - ar) + ac)) in = !in;
                        prev = cur;
                    }
                    inside = in;
                }
            for (int64_t ch = 0; ch < img.channels; ++ch) {
                float &v = img.reference(r, c, ch);
                if (inside) {
                    if (ud->f_bounded) ud->f_bounded(r, c, ch, std::ref(img), std::ref(mask_img), v);
                } else {
                    if (ud->f_unbounded) ud->f_unbounded(r, c, ch, std::ref(img), std::ref(mask_img), v);
                }
                if (ud->f_visitor) ud->f_visitor(r, c, ch, std::ref(img), std::ref(mask_img), v);
            }
        }
    return true;
}
