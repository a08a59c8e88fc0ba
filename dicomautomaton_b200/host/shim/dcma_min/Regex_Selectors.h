// Stand-in for the part of the reference's src/Regex_Selectors.h that RegisterImagesDemons.cc uses: Compile_Regex,
// All_IAs, Whitelist for image arrays (positional selectors only: first, last, all, none, #N, #-N, and !negation) and
// IAWhitelistOpArgDoc.  NOT a product component.
#pragma once
#include <iterator>
#include <list>
#include <memory>
#include <regex>
#include <stdexcept>
#include <string>

#include "Structs.h"

inline std::regex Compile_Regex(const std::string &s) {
    return std::regex(s, std::regex::icase | std::regex::nosubs | std::regex::optimize | std::regex::extended);
}

using IA_iter_list = std::list<std::list<std::shared_ptr<Image_Array>>::iterator>;

inline IA_iter_list All_IAs(Drover &d) {
    IA_iter_list out;
    for (auto it = d.image_data.begin(); it != d.image_data.end(); ++it) out.push_back(it);
    return out;
}

inline IA_iter_list Whitelist(IA_iter_list all, std::string sel) {
    bool negate = false;
    if (!sel.empty() && sel.front() == '!') {
        negate = true;
        sel.erase(sel.begin());
    }
    IA_iter_list picked;
    const auto n = static_cast<long>(all.size());
    auto nth = [&](long i) {
        if (i >= 0 && i < n) picked.push_back(*std::next(all.begin(), i));
    };
    if (std::regex_match(sel, Compile_Regex("^al?l?$"))) picked = all;
    else if (std::regex_match(sel, Compile_Regex("^no?n?e?$"))) {}
    else if (std::regex_match(sel, Compile_Regex("^fi?r?s?t?$"))) nth(0);
    else if (std::regex_match(sel, Compile_Regex("^la?s?t?$"))) nth(n - 1);
    else if (sel.size() > 2 && sel[0] == '#' && sel[1] == '-') nth(n - 1 - std::stol(sel.substr(2)));
    else if (sel.size() > 1 && sel[0] == '#') nth(std::stol(sel.substr(1)));
    else throw std::invalid_argument("selector '" + sel + "' is not supported by this stand-in");
    if (!negate) return picked;
    IA_iter_list rest;
    for (const auto &it : all)
        if (std::find(picked.begin(), picked.end(), it) == picked.end()) rest.push_back(it);
    return rest;
}

inline OperationArgDoc IAWhitelistOpArgDoc() {
    OperationArgDoc out;
    out.name = "ImageSelection";
    out.desc = "Select one or more image arrays.";
    out.default_val = "all";
    out.expected = true;
    out.examples = {"last", "first", "all", "none", "#0", "#-0", "!last"};
    return out;
}
