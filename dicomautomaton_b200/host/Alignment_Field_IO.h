// Alignment_Field_IO.h -- the text (de)serialisation of a deformation field with the numbers converted on a B200.
//
// Bodies for `deformation_field::write_to(std::ostream&)` and `deformation_field::read_from(std::istream&)` of the
// reference (src/Alignment_Field.h:75-76, src/Alignment_Field.cc:322-436): same file format byte for byte, same return
// values and warnings.  The per-image header lines are written/read on the host exactly as the reference does; the
// rows*columns*3 numbers of every image -- all of the volume of a `.trans` file -- go through
// dcma_b200_format_doubles / dcma_b200_parse_doubles (include/dcma_b200/demons_c_api.h).  There is no CPU fallback for
// the numbers: without a device both functions warn and return false.
//
// Install ("Field export"): in src/Alignment_Field.cc replace the two member bodies by
//     bool deformation_field::write_to(std::ostream &os) const { return dcma_b200_io::write_field_text(this->field, os); }
//     bool deformation_field::read_from(std::istream &is) {
//         planar_image_collection<double, double> f;
//         if(!dcma_b200_io::read_field_text(is, f)) return false;
//         this->swap_and_rebuild(f);  return true; }
#pragma once
#include <iosfwd>

#include "YgorImages.h"

namespace dcma_b200_io {

// device < 0: the current CUDA device.
bool write_field_text(const planar_image_collection<double, double> &field, std::ostream &os, int device = -1);

// Reads the REST of `is` into memory (a field is the last item of a `.trans` file, src/Transformation_File_Loader.cc:
// 142-149), so unlike the reference the stream is left at its end, not just after the last number.
bool read_field_text(std::istream &is, planar_image_collection<double, double> &field, int device = -1);

}  // namespace dcma_b200_io
