// Alignment_Field_IO.cc -- see the header.  Follows src/Alignment_Field.cc:322-436 statement by statement on the host
// side; only the two `for(val : img.data)` loops (:346-348 and :418-420) are replaced by one GPU call each.
#include "Alignment_Field_IO.h"

#include <cstdint>
#include <istream>
#include <iterator>
#include <limits>
#include <ostream>
#include <streambuf>
#include <string>
#include <vector>

#include "YgorBase64.h"
#include "YgorLog.h"
#include "YgorMath.h"
#include "YgorMisc.h"

#include "dcma_b200/demons_c_api.h"

namespace {
// An istream over bytes already in memory, whose read position can be queried and moved.
struct span_buf : public std::streambuf {
    span_buf(char *b, char *e) { setg(b, b, e); }
    size_t pos() const { return static_cast<size_t>(gptr() - eback()); }
    void seek(size_t p) { setg(eback(), eback() + p, egptr()); }
};
}  // namespace

bool dcma_b200_io::write_field_text(const planar_image_collection<double, double> &field, std::ostream &os, int device) {
    const auto original_precision = os.precision();
    os.precision(std::numeric_limits<double>::max_digits10);

    const auto N_imgs = static_cast<int64_t>(field.images.size());
    os << N_imgs << '\n';

    std::vector<char> text;
    bool ok = true;
    for (const auto &img : field.images) {
        os << img.rows << " " << img.columns << " " << img.channels << '\n';
        os << img.pxl_dx << " " << img.pxl_dy << " " << img.pxl_dz << '\n';
        os << img.anchor << '\n';
        os << img.offset << '\n';
        os << img.row_unit << '\n';
        os << img.col_unit << '\n';

        os << "num_metadata= " << img.metadata.size() << '\n';
        for (const auto &mp : img.metadata) os << Base64::EncodeFromString(mp.first) << " " << Base64::EncodeFromString(mp.second) << '\n';

        // pixel data: one "%.17g" line per value, converted on the GPU
        const auto n = static_cast<int64_t>(img.data.size());
        text.resize(static_cast<size_t>(n) * 25u);
        int64_t len = 0;
        char err[512] = {0};
        if (dcma_b200_format_doubles(img.data.data(), n, text.data(), static_cast<int64_t>(text.size()), &len, device, err, sizeof err) != DCMA_B200_OK) {
            YLOGWARN("Unable to write deformation field: " << err);
            ok = false;
            break;
        }
        os.write(text.data(), static_cast<std::streamsize>(len));
    }

    os.precision(original_precision);
    os.flush();
    return ok && !os.fail();
}

bool dcma_b200_io::read_field_text(std::istream &is_in, planar_image_collection<double, double> &field, int device) {
    std::string all((std::istreambuf_iterator<char>(is_in)), std::istreambuf_iterator<char>());
    span_buf sb(all.data(), all.data() + all.size());
    std::istream is(&sb);

    int64_t N_imgs = 0;
    is >> N_imgs;
    if (is.fail() || !isininc(1, N_imgs, 10'000)) {
        YLOGWARN("Number of images could not be read, or is invalid.");
        return false;
    }

    planar_image_collection<double, double> new_field;
    try {
        for (int64_t i = 0; i < N_imgs; ++i) {
            int64_t rows = 0, cols = 0, channels = 0;
            is >> rows >> cols >> channels;
            if (is.fail() || rows <= 0 || cols <= 0 || channels != 3) {
                YLOGWARN("Image dimensions could not be read, or are invalid.");
                return false;
            }
            double pxl_dx = 0.0, pxl_dy = 0.0, pxl_dz = 0.0;
            is >> pxl_dx >> pxl_dy >> pxl_dz;
            if (is.fail()) {
                YLOGWARN("Pixel dimensions could not be read.");
                return false;
            }
            vec3<double> anchor, offset, row_unit, col_unit;
            is >> anchor >> offset >> row_unit >> col_unit;
            if (is.fail()) {
                YLOGWARN("Image spatial parameters could not be read.");
                return false;
            }

            planar_image<double, double> img;
            img.init_orientation(row_unit, col_unit);
            img.init_buffer(rows, cols, channels);
            img.init_spatial(pxl_dx, pxl_dy, pxl_dz, anchor, offset);

            std::string label;
            int64_t N_metadata = 0;
            is >> label;  // 'num_metadata='
            is >> N_metadata;
            if (is.fail() || N_metadata < 0) {
                YLOGWARN("Metadata count could not be read.");
                return false;
            }
            for (int64_t m = 0; m < N_metadata; ++m) {
                std::string enc_key, enc_val;
                is >> enc_key >> enc_val;
                if (is.fail()) {
                    YLOGWARN("Metadata entry could not be read.");
                    return false;
                }
                img.metadata[Base64::DecodeToString(enc_key)] = Base64::DecodeToString(enc_val);
            }

            // pixel data: rows*cols*3 stream extractions, done on the GPU from the bytes at the read position
            const size_t at = sb.pos();
            int64_t consumed = 0;
            char err[512] = {0};
            if (dcma_b200_parse_doubles(all.data() + at, static_cast<int64_t>(all.size() - at), static_cast<int64_t>(img.data.size()),
                                        img.data.data(), &consumed, device, err, sizeof err) != DCMA_B200_OK) {
                YLOGWARN("Pixel data could not be read. (" << err << ")");
                return false;
            }
            sb.seek(at + static_cast<size_t>(consumed));

            new_field.images.push_back(img);
        }
    } catch (const std::exception &e) {
        YLOGWARN("Failed to read deformation field: " << e.what());
        return false;
    }

    field.Swap(new_field);
    return !is.fail();
}
