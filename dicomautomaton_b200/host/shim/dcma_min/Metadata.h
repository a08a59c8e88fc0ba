// Stand-in for the part of the reference's src/Metadata.h that WarpImages.cc uses.  NOT a product component.
#pragma once
#include <map>
#include <optional>
#include <string>

using metadata_map_t = std::map<std::string, std::string>;
enum class meta_evolve { none, iterate };

// The reference fills in the keys a basic image needs and, with `iterate`, advances the per-image ones.
inline metadata_map_t coalesce_metadata_for_basic_image(const metadata_map_t &ref, meta_evolve e = meta_evolve::none) {
    metadata_map_t out = ref;
    out.emplace("FrameOfReferenceUID", "1.2.826.0.1.3680043.stand.in");
    out.emplace("Modality", "OT");
    if (e == meta_evolve::iterate) {
        auto &n = out["InstanceNumber"];
        n = std::to_string((n.empty() ? 0 : std::stol(n)) + 1);
    }
    return out;
}

template <class T>
std::optional<T> get_as(const metadata_map_t &m, const std::string &key);
template <>
inline std::optional<std::string> get_as<std::string>(const metadata_map_t &m, const std::string &key) {
    const auto it = m.find(key);
    if (it == m.end()) return std::nullopt;
    return it->second;
}
