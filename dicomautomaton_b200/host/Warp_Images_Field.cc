// Warp_Images_Field.cc -- see the header.
#include "Warp_Images_Field.h"

#include <stdexcept>

#include "Dense_Stack.h"

bool dcma_b200_host::warp_stack_with_field(const deformation_field &field, const planar_image_collection<float, double> &source,
                                           planar_image_collection<float, double> &target,
                                           const std::vector<unsigned char> &bounded, int64_t channel, float inaccessible_value,
                                           int device) {
    using namespace dcma_b200_host::detail;
    if (source.images.empty() || target.images.empty()) return false;
    const auto &fcoll = field.get_imagecoll_crefw().get();
    dcma_b200_grid fg, sg, og;
    if (!grid_of(fcoll, fg) || fg.geom.channels != 3 || !grid_of(source, sg) || !grid_of(target, og)) return false;
    if (channel >= og.geom.channels || (channel < 0 ? og.geom.channels : channel + 1) > sg.geom.channels) return false;
    const size_t nf = static_cast<size_t>(fg.geom.slices * fg.geom.rows * fg.geom.cols * 3);
    const size_t ns = static_cast<size_t>(sg.geom.slices * sg.geom.rows * sg.geom.cols * sg.geom.channels);
    const size_t nvox = static_cast<size_t>(og.geom.slices * og.geom.rows * og.geom.cols);
    if (!bounded.empty() && bounded.size() != nvox) throw std::invalid_argument("warp_stack_with_field: one flag per target voxel expected");
    Pinned<double> fld(nf);
    Pinned<float> src(ns), out(nvox * static_cast<size_t>(og.geom.channels));
    marshal_into(fcoll, fld.data());
    marshal_into(source, src.data());
    marshal_into(target, out.data());
    char err[512] = {0};
    check_rc(dcma_b200_warp_to_grid(&fg, fld.data(), &sg, src.data(), &og, bounded.empty() ? nullptr : bounded.data(),
                                    static_cast<int32_t>(channel), inaccessible_value, out.data(), device, err, sizeof err), err);
    fill_from_dense(target, out.data(), og.geom.rows, og.geom.cols, og.geom.channels);
    return true;
}
