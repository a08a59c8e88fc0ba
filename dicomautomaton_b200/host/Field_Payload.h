// Field_Payload.h -- the NUMBERS of a serialised deformation field, converted on a B200.
//
// A stored field (`deformation_field::write_to` / `read_from`, reference src/Alignment_Field.cc:322-436) is a few header
// lines per image followed by rows*columns*3 doubles, one per line at max_digits10: at 512x512x256 that is 2.0e8 numbers
// and ~4.8 GB of text, minutes of iostream time.  The header lines stay in the reference's own code; this codec replaces
// only the two per-value loops (:346-348 `os << val << '\n'` and :418-420 `is >> val`) and knows nothing about images.
// patches/Alignment_Field.patch is the two-hunk change that calls it.
//
// Same bytes as the iostream loop writes, same doubles as it reads (csrc/text_format.cuh); no CPU fallback: without a
// device both functions log a warning and report failure through the stream state, like any other I/O error.
#pragma once
#include <cstdint>
#include <iosfwd>

namespace dcma_b200_io {

// Writes n values as n lines of printf("%.17g\n").  Sets failbit on `os` and returns false on error.  device < 0: current.
bool write_payload(std::ostream &os, const double *values, int64_t n, int device = -1);

// Performs the n extractions `is >> out[i]`: leading white space skipped, the stream is left just behind the last number.
// Sets failbit on `is` and returns false when fewer than n numbers could be read (out[] is then unspecified).
// The stream must be seekable (files, string streams): the codec reads ahead in bounded blocks and seeks back.
bool read_payload(std::istream &is, double *out, int64_t n, int device = -1);

}  // namespace dcma_b200_io
