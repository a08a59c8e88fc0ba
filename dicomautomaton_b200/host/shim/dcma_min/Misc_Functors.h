// Stand-in for src/YgorImages_Functors/Grouping/Misc_Functors.h: the grouping tag WarpImages.cc passes.  NOT a product
// component.
#pragma once
struct GroupIndividualImagesTag {};
inline constexpr GroupIndividualImagesTag GroupIndividualImages{};
