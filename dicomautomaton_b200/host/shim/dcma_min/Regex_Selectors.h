// Stand-in for the part of the reference's src/Regex_Selectors.h that RegisterImagesDemons.cc and WarpImages.cc use:
// Compile_Regex; All_IAs / All_T3s with Whitelist (positional selectors only: first, last, all, none, #N, #-N, and
// !negation); All_CCs with Whitelist by ROI label regexes; the matching *WhitelistOpArgDoc helpers.  NOT a product component.
#pragma once
#include <functional>
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

template <class IterList>
inline IterList Whitelist_positional(IterList all, std::string sel) {
    bool negate = false;
    if (!sel.empty() && sel.front() == '!') {
        negate = true;
        sel.erase(sel.begin());
    }
    IterList picked;
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
    IterList rest;
    for (const auto &it : all)
        if (std::find(picked.begin(), picked.end(), it) == picked.end()) rest.push_back(it);
    return rest;
}

inline IA_iter_list Whitelist(IA_iter_list all, std::string sel) { return Whitelist_positional(std::move(all), std::move(sel)); }

using T3_iter_list = std::list<std::list<std::shared_ptr<Transform3>>::iterator>;
inline T3_iter_list All_T3s(Drover &d) {
    T3_iter_list out;
    for (auto it = d.trans_data.begin(); it != d.trans_data.end(); ++it) out.push_back(it);
    return out;
}
inline T3_iter_list Whitelist(T3_iter_list all, std::string sel) { return Whitelist_positional(std::move(all), std::move(sel)); }

using CC_ref_list = std::list<std::reference_wrapper<contour_collection<double>>>;
inline CC_ref_list All_CCs(Drover &d) {
    CC_ref_list out;
    if (d.contour_data)
        for (auto &cc : d.contour_data->ccs) out.push_back(std::ref(cc));
    return out;
}
// collections whose first contour's "ROIName" and "NormalizedROIName" match the two regexes; the third selector is positional
inline CC_ref_list Whitelist(CC_ref_list all, const std::string &roi_regex, const std::string &norm_regex, const std::string &sel = "all") {
    const auto r1 = Compile_Regex(roi_regex), r2 = Compile_Regex(norm_regex);
    CC_ref_list out;
    for (auto &cc : all) {
        if (cc.get().contours.empty()) continue;
        const auto &md = cc.get().contours.front().metadata;
        const auto a = md.find("ROIName"), b = md.find("NormalizedROIName");
        if (a == md.end() || b == md.end()) continue;
        if (std::regex_match(a->second, r1) && std::regex_match(b->second, r2)) out.push_back(cc);
    }
    if (!std::regex_match(sel, Compile_Regex("^al?l?$"))) throw std::invalid_argument("ROISelection '" + sel + "' is not supported by this stand-in");
    return out;
}

inline OperationArgDoc simple_selector_doc(const char *name, const char *desc, const char *def) {
    OperationArgDoc out;
    out.name = name;
    out.desc = desc;
    out.default_val = def;
    out.expected = true;
    out.examples = {def};
    return out;
}
inline OperationArgDoc T3WhitelistOpArgDoc() { return simple_selector_doc("TransformSelection", "Select one or more transform objects.", "last"); }
inline OperationArgDoc NCWhitelistOpArgDoc() { return simple_selector_doc("NormalizedROILabelRegex", "Regex of normalised ROI labels.", ".*"); }
inline OperationArgDoc RCWhitelistOpArgDoc() { return simple_selector_doc("ROILabelRegex", "Regex of ROI labels.", ".*"); }
inline OperationArgDoc CCWhitelistOpArgDoc() { return simple_selector_doc("ROISelection", "Select one or more contour collections.", "all"); }

inline OperationArgDoc IAWhitelistOpArgDoc() {
    OperationArgDoc out;
    out.name = "ImageSelection";
    out.desc = "Select one or more image arrays.";
    out.default_val = "all";
    out.expected = true;
    out.examples = {"last", "first", "all", "none", "#0", "#-0", "!last"};
    return out;
}
