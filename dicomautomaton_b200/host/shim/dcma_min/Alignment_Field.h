// Stand-in for the reference's src/Alignment_Field.h, used ONLY to build the host-side tests in this container, where
// the real file cannot compile (it needs Ygor, Structs.h, Boost...).  Inside the reference tree the drop-in uses the
// reference's own deformation_field unchanged.  Mirrors src/Alignment_Field.h:34-77 / .cc:45-153 for the members the
// Demons path touches: construction from a 3-channel regular stack, the two accessors and transform().  The text
// (de)serialisation members write_to / read_from are only DECLARED here: their definitions are the reference's own
// (src/Alignment_Field.cc:322-436) with patches/Alignment_Field.patch applied, compiled in place by
// host/build_host_tests.py::build_field_io -- nothing of them lives in this repository.
#pragma once
#include <functional>
#include <limits>
#include <optional>
#include <stdexcept>

#include <istream>
#include <ostream>

#include "YgorImages.h"
#include "YgorMath.h"

class deformation_field {
    planar_image_collection<double, double> field;
    std::optional<planar_image_adjacency<double, double>> adj;
    void rebuild() {
        const auto unit = field.images.front().ortho_unit();
        adj.emplace(std::list<std::reference_wrapper<planar_image<double, double>>>{},
                    std::list<std::reference_wrapper<planar_image_collection<double, double>>>{std::ref(field)}, unit);
    }
  public:
    deformation_field() = delete;
    deformation_field(planar_image_collection<double, double> &&in) {  // .cc:51-53, 93-123
        std::list<std::reference_wrapper<planar_image<double, double>>> sel;
        for (auto &img : in.images) {
            if (img.channels != 3) throw std::invalid_argument("Encountered image without three channels");
            sel.push_back(std::ref(img));
        }
        if (sel.empty()) throw std::invalid_argument("No images provided");
        if (!Images_Form_Regular_Grid(sel)) throw std::invalid_argument("Images do not form a rectilinear grid. Cannot continue");
        field.Swap(in);
        rebuild();
    }
    deformation_field(std::istream &is) {  // .cc:45-49: defers to read_from(), throws on errors
        if (!read_from(is)) throw std::invalid_argument("Unable to read deformation field from stream");
    }
    bool write_to(std::ostream &os) const;  // .cc:322-357, defined by the (patched) reference source
    bool read_from(std::istream &is);       // .cc:359-436, likewise
    void swap_and_rebuild(planar_image_collection<double, double> &in) {  // .cc:93-123 (validation as in the constructor)
        deformation_field tmp(std::move(in));
        field.Swap(tmp.field);
        rebuild();
    }
    deformation_field(const deformation_field &o) : field(o.field) { rebuild(); }
    deformation_field(deformation_field &&o) : field(std::move(o.field)) { rebuild(); }
    deformation_field &operator=(const deformation_field &o) { if (this != &o) { field = o.field; rebuild(); } return *this; }
    deformation_field &operator=(deformation_field &&o) { if (this != &o) { field = std::move(o.field); rebuild(); } return *this; }

    std::reference_wrapper<const planar_image_collection<double, double>> get_imagecoll_crefw() const { return std::cref(field); }
    std::reference_wrapper<const planar_image_adjacency<double, double>> get_adjacency_crefw() const { return std::cref(adj.value()); }
    void apply_to(vec3<double> &v) const { v = transform(v); }  // .cc:161-165
    vec3<double> transform(const vec3<double> &v) const {  // .cc:138-153
        const double oob = std::numeric_limits<double>::quiet_NaN();
        return v + vec3<double>(adj->trilinearly_interpolate(v, 0, oob), adj->trilinearly_interpolate(v, 1, oob),
                                adj->trilinearly_interpolate(v, 2, oob));
    }
};
