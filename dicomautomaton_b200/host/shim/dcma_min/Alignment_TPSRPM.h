// Stand-in for the reference's Alignment_TPSRPM.h: the type only (WarpImages.cc refuses thin-plate splines).  NOT a product
// component.
#pragma once
class thin_plate_spline {};
