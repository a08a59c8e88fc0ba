// Field_Payload.cc -- see the header.
#include "Field_Payload.h"

#include <algorithm>
#include <istream>
#include <ostream>
#include <vector>

#include "YgorLog.h"

#include "dcma_b200/demons_c_api.h"

namespace {
constexpr int64_t kLineMax = 25;            // "-d.dddddddddddddddde-ddd\n": the longest %.17g line
constexpr int64_t kBatch = 8ll << 20;       // numbers per conversion call (the library's own pass size)
}  // namespace

bool dcma_b200_io::write_payload(std::ostream &os, const double *values, int64_t n, int device) {
    std::vector<char> text;
    for (int64_t at = 0; at < n; at += kBatch) {
        const int64_t m = std::min(kBatch, n - at);
        text.resize(static_cast<size_t>(m * kLineMax));
        int64_t len = 0;
        char err[512] = {0};
        if (dcma_b200_format_doubles(values + at, m, text.data(), static_cast<int64_t>(text.size()), &len, device, err, sizeof err) != DCMA_B200_OK) {
            YLOGWARN("Unable to convert field values to text on the GPU: " << err);
            os.setstate(std::ios::failbit);
            return false;
        }
        os.write(text.data(), static_cast<std::streamsize>(len));
    }
    return !os.fail();
}

bool dcma_b200_io::read_payload(std::istream &is, double *out, int64_t n, int device) {
    if (n <= 0) return !is.fail();
    if (is.fail()) return false;
    const std::istream::pos_type start = is.tellg();
    if (start == std::istream::pos_type(-1)) {
        YLOGWARN("Field payload: the stream is not seekable");
        is.setstate(std::ios::failbit);
        return false;
    }
    // Read ahead no more than the text n numbers can plausibly occupy; a hand-edited file with longer tokens or more
    // white space only costs another, larger attempt.
    std::vector<char> text;
    for (int64_t cap = n * (kLineMax + 1) + 64;; cap *= 4) {
        text.resize(static_cast<size_t>(cap));
        is.clear();
        is.seekg(start);
        is.read(text.data(), static_cast<std::streamsize>(cap));
        const int64_t got = static_cast<int64_t>(is.gcount());
        const bool whole_rest = got < cap;  // hit the end of the stream
        is.clear();
        int64_t consumed = 0;
        char err[512] = {0};
        const int rc = dcma_b200_parse_doubles(text.data(), got, n, out, &consumed, device, err, sizeof err);
        // a block cut in the middle of the last wanted token would truncate that number: accept only when at least one
        // byte follows the consumed range, or the block is the whole rest of the stream
        if (rc == DCMA_B200_OK && (consumed < got || whole_rest)) {
            is.seekg(start + static_cast<std::streamoff>(consumed));
            return true;
        }
        // only "too few / malformed numbers in this block" (EINVAL) can be cured by reading further ahead
        if (whole_rest || (rc != DCMA_B200_OK && rc != DCMA_B200_EINVAL)) {
            YLOGWARN("Field payload could not be read: " << err);
            is.seekg(start);
            is.setstate(std::ios::failbit);
            return false;
        }
    }
}
