// Minimal stand-in for Ygor's YgorString.h: the two helpers RegisterImagesDemons.cc uses (the "..."_s literal and
// SplitStringToVector with the 'd' = drop-the-delimiter mode).  NOT a product component.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

inline std::string operator""_s(const char *s, std::size_t n) { return std::string(s, n); }

inline std::vector<std::string> SplitStringToVector(const std::string &in, char delim, char /*mode: 'd' drops the delimiter*/) {
    std::vector<std::string> out;
    std::string cur;
    for (char c : in) {
        if (c == delim) {
            out.push_back(cur);
            cur.clear();
        } else {
            cur += c;
        }
    }
    out.push_back(cur);
    return out;
}
