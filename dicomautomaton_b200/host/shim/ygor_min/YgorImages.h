// Minimal stand-in for Ygor's YgorImages.h: planar_image, planar_image_collection, planar_image_adjacency and
// Images_Form_Regular_Grid -- only the members the Demons path and its tests use (last row).
// Ygor is absent here (not vendored, no network), so the interpolation rules below are OUR reading of how the reference
// uses them; they are pinned only by the reference's own tests (src/Alignment_Demons_Tests.cc:171-202, 276-299,
// 379-405; src/Alignment_Field_Tests.cc:144-193, 459-478).  "parity unpinned" beyond those.
//
// Geometry convention (src/Alignment_Demons_Tests.cc:45-118, src/Alignment_Field.cc:282-283):
//   position(row, col) = anchor + offset + row_unit * (pxl_dx * col) + col_unit * (pxl_dy * row)
//   x <-> column <-> row_unit <-> pxl_dx ; y <-> row <-> col_unit <-> pxl_dy ; z <-> slice <-> pxl_dz
#pragma once
#include <algorithm>
#include <any>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <list>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "YgorMath.h"

template <class T, class R>
class planar_image {
  public:
    int64_t rows = 0, columns = 0, channels = 0;
    R pxl_dx = R(1), pxl_dy = R(1), pxl_dz = R(1);
    vec3<R> anchor, offset, row_unit = vec3<R>(1, 0, 0), col_unit = vec3<R>(0, 1, 0);
    std::vector<T> data;
    std::map<std::string, std::string> metadata;

    void init_orientation(const vec3<R> &r, const vec3<R> &c) { row_unit = r; col_unit = c; }
    void init_buffer(int64_t r, int64_t c, int64_t ch) {
        if (r <= 0 || c <= 0 || ch <= 0) throw std::invalid_argument("bad image extent");
        rows = r; columns = c; channels = ch;
        data.assign(static_cast<size_t>(r * c * ch), T(0));
    }
    void fill_pixels(T v) { std::fill(data.begin(), data.end(), v); }
    void init_spatial(R dx, R dy, R dz, const vec3<R> &a, const vec3<R> &o) { pxl_dx = dx; pxl_dy = dy; pxl_dz = dz; anchor = a; offset = o; }
    int64_t index(int64_t r, int64_t c, int64_t ch = 0) const {
        if (r < 0 || r >= rows || c < 0 || c >= columns || ch < 0 || ch >= channels) return -1;
        return (r * columns + c) * channels + ch;
    }
    T value(int64_t r, int64_t c, int64_t ch) const {
        const auto i = index(r, c, ch);
        if (i < 0) throw std::runtime_error("planar_image::value out of bounds");
        return data[static_cast<size_t>(i)];
    }
    T &reference(int64_t r, int64_t c, int64_t ch) {
        const auto i = index(r, c, ch);
        if (i < 0) throw std::runtime_error("planar_image::reference out of bounds");
        return data[static_cast<size_t>(i)];
    }
    vec3<R> position(int64_t r, int64_t c) const {
        return anchor + offset + row_unit * (pxl_dx * static_cast<R>(c)) + col_unit * (pxl_dy * static_cast<R>(r));
    }
    vec3<R> ortho_unit() const { return row_unit.Cross(col_unit).unit(); }
    vec3<R> center() const { return position(0, 0) + (position(rows - 1, columns - 1) - position(0, 0)) * R(0.5); }

    // Fractional (row, col) of a point projected onto the image plane, and its signed distance from the plane.
    void fractional(const vec3<R> &p, R &fr, R &fc, R &dist) const {
        const vec3<R> d = p - (anchor + offset);
        fc = d.Dot(row_unit.unit()) / pxl_dx;
        fr = d.Dot(col_unit.unit()) / pxl_dy;
        dist = d.Dot(ortho_unit());
    }
    // In-plane bilinear sample; a pixel is an AREA, so up to half a pixel beyond the outermost centres is inside and
    // sampled clamped to the edge.  Returns false if outside.
    bool bilinear(R fr, R fc, int64_t ch, R &out) const {
        if (!(std::isfinite(fr) && std::isfinite(fc))) return false;
        if (fr < R(-0.5) || fr > static_cast<R>(rows) - R(0.5) || fc < R(-0.5) || fc > static_cast<R>(columns) - R(0.5)) return false;
        fr = std::min(std::max(fr, R(0)), static_cast<R>(rows - 1));
        fc = std::min(std::max(fc, R(0)), static_cast<R>(columns - 1));
        const int64_t r0 = std::min<int64_t>(static_cast<int64_t>(std::floor(fr)), std::max<int64_t>(rows - 2, 0));
        const int64_t c0 = std::min<int64_t>(static_cast<int64_t>(std::floor(fc)), std::max<int64_t>(columns - 2, 0));
        const int64_t r1 = std::min<int64_t>(r0 + 1, rows - 1), c1 = std::min<int64_t>(c0 + 1, columns - 1);
        const R tr = fr - static_cast<R>(r0), tc = fc - static_cast<R>(c0);
        const R v00 = static_cast<R>(value(r0, c0, ch)), v01 = static_cast<R>(value(r0, c1, ch));
        const R v10 = static_cast<R>(value(r1, c0, ch)), v11 = static_cast<R>(value(r1, c1, ch));
        out = (v00 * (R(1) - tc) + v01 * tc) * (R(1) - tr) + (v10 * (R(1) - tc) + v11 * tc) * tr;
        return true;
    }
    bool nearest(R fr, R fc, int64_t ch, R &out) const {
        if (!(std::isfinite(fr) && std::isfinite(fc))) return false;
        const int64_t r = static_cast<int64_t>(std::floor(fr + R(0.5))), c = static_cast<int64_t>(std::floor(fc + R(0.5)));
        if (r < 0 || r >= rows || c < 0 || c >= columns) return false;
        out = static_cast<R>(value(r, c, ch));
        return true;
    }
};

namespace ygor_min {
// Images sorted along `unit`, with the machinery shared by the collection and the adjacency index.
template <class T, class R>
struct stack_index {
    std::vector<const planar_image<T, R> *> imgs;  // sorted by position along the stacking direction
    std::vector<R> pos;
    void build(const std::list<planar_image<T, R>> &l, const vec3<R> &unit) {
        imgs.clear(); pos.clear();
        std::vector<std::pair<R, const planar_image<T, R> *>> v;
        for (const auto &im : l) v.emplace_back((im.anchor + im.offset).Dot(unit), &im);
        std::stable_sort(v.begin(), v.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
        for (auto &e : v) { pos.push_back(e.first); imgs.push_back(e.second); }
    }
    // trilinear: bilinear in the two bracketing planes, linear across; one plane (or beyond the outermost centres but
    // inside the slab thickness) -> that plane alone.  single_plane_nearest selects the collection's degraded mode.
    T sample(const vec3<R> &p, const vec3<R> &unit, int64_t ch, T oob, bool single_plane_nearest) const {
        if (imgs.empty() || !p.isfinite()) return oob;
        const R s = p.Dot(unit);
        const size_t n = imgs.size();
        auto in_plane = [&](size_t i, R &out) {
            R fr, fc, d;
            imgs[i]->fractional(p, fr, fc, d);
            return (n == 1 && single_plane_nearest) ? imgs[i]->nearest(fr, fc, ch, out) : imgs[i]->bilinear(fr, fc, ch, out);
        };
        const R half_lo = imgs.front()->pxl_dz * R(0.5), half_hi = imgs.back()->pxl_dz * R(0.5);
        if (s < pos.front() - half_lo || s > pos.back() + half_hi) return oob;
        R out;
        if (n == 1 || s <= pos.front()) return in_plane(0, out) ? static_cast<T>(out) : oob;
        if (s >= pos.back()) return in_plane(n - 1, out) ? static_cast<T>(out) : oob;
        const size_t hi = static_cast<size_t>(std::upper_bound(pos.begin(), pos.end(), s) - pos.begin());
        const size_t lo = hi - 1;
        R a, b;
        if (!in_plane(lo, a) || !in_plane(hi, b)) return oob;
        const R t = (s - pos[lo]) / (pos[hi] - pos[lo]);
        return static_cast<T>(a * (R(1) - t) + b * t);
    }
};
}  // namespace ygor_min

template <class T, class R>
class planar_image_collection {
  public:
    std::list<planar_image<T, R>> images;
    void Swap(planar_image_collection &o) { images.swap(o.images); }
    // Degrades to nearest-neighbour in-plane when only one image plane is present (as the reference's comment at
    // src/Alignment_Demons.cc:905-908 describes Ygor's behaviour).
    T trilinearly_interpolate(const vec3<R> &p, int64_t ch, T oob) const {
        if (images.empty()) return oob;
        const vec3<R> unit = images.front().ortho_unit();
        ygor_min::stack_index<T, R> idx;
        idx.build(images, unit);
        return idx.sample(p, unit, ch, oob, true);
    }
    using images_list_it_t = typename std::list<planar_image<T, R>>::iterator;
    // one call of `op` per image (the only grouping the Demons-path operations use: GroupIndividualImages); serial here
    template <class Grouper, class Op>
    bool Process_Images_Parallel(Grouper, Op op, std::list<std::reference_wrapper<planar_image_collection<T, R>>> external,
                                 std::list<std::reference_wrapper<contour_collection<R>>> ccs, std::any user_data = {}) {
        for (auto it = images.begin(); it != images.end(); ++it)
            if (!op(it, std::list<images_list_it_t>{it}, external, ccs, user_data)) return false;
        return true;
    }
    std::map<std::string, std::string> get_common_metadata(const std::list<images_list_it_t> & = {}) const {
        std::map<std::string, std::string> out;
        if (images.empty()) return out;
        out = images.front().metadata;
        for (const auto &im : images)
            for (auto it = out.begin(); it != out.end();) {
                auto f = im.metadata.find(it->first);
                if (f == im.metadata.end() || f->second != it->second) it = out.erase(it); else ++it;
            }
        return out;
    }
};

template <class T, class R>
class planar_image_adjacency {
    ygor_min::stack_index<T, R> idx_;
    vec3<R> unit_;
  public:
    planar_image_adjacency(const std::list<std::reference_wrapper<planar_image<T, R>>> &imgs,
                           const std::list<std::reference_wrapper<planar_image_collection<T, R>>> &colls,
                           const vec3<R> &unit) : unit_(unit.unit()) {
        std::list<planar_image<T, R>> dummy;
        std::vector<std::pair<R, const planar_image<T, R> *>> v;
        for (auto &r : imgs) v.emplace_back((r.get().anchor + r.get().offset).Dot(unit_), &r.get());
        for (auto &c : colls) for (auto &im : c.get().images) v.emplace_back((im.anchor + im.offset).Dot(unit_), &im);
        std::stable_sort(v.begin(), v.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
        for (auto &e : v) { idx_.pos.push_back(e.first); idx_.imgs.push_back(e.second); }
    }
    T trilinearly_interpolate(const vec3<R> &p, int64_t ch, T oob) const { return idx_.sample(p, unit_, ch, oob, false); }
    const std::vector<const planar_image<T, R> *> &int_to_img = idx_.imgs;  // index along the stacking direction -> image
};

// Parallel planes with the same in-plane grid (spacing along the stacking direction may vary).
template <class T, class R>
bool Images_Form_Rectilinear_Grid(const std::list<std::reference_wrapper<planar_image<T, R>>> &imgs, R eps = R(1e-6)) {
    if (imgs.empty()) return false;
    const auto &a = imgs.front().get();
    const vec3<R> unit = a.ortho_unit();
    for (const auto &r : imgs) {
        const auto &b = r.get();
        if (b.rows != a.rows || b.columns != a.columns) return false;
        if (std::abs(b.pxl_dx - a.pxl_dx) > eps || std::abs(b.pxl_dy - a.pxl_dy) > eps) return false;
        if ((b.row_unit - a.row_unit).length() > eps || (b.col_unit - a.col_unit).length() > eps) return false;
        const vec3<R> d = (b.anchor + b.offset) - (a.anchor + a.offset);
        if ((d - unit * d.Dot(unit)).length() > eps * std::max<R>(R(1), d.length())) return false;
    }
    return true;
}

// Same extents, spacing and orientation everywhere, and equally spaced along the stacking direction.
template <class T, class R>
bool Images_Form_Regular_Grid(const std::list<std::reference_wrapper<planar_image<T, R>>> &imgs, R eps = R(1e-6)) {
    if (imgs.empty()) return false;
    const auto &a = imgs.front().get();
    const vec3<R> unit = a.ortho_unit();
    std::vector<R> pos;
    for (const auto &r : imgs) {
        const auto &b = r.get();
        if (b.rows != a.rows || b.columns != a.columns) return false;
        if (std::abs(b.pxl_dx - a.pxl_dx) > eps || std::abs(b.pxl_dy - a.pxl_dy) > eps || std::abs(b.pxl_dz - a.pxl_dz) > eps) return false;
        if ((b.row_unit - a.row_unit).length() > eps || (b.col_unit - a.col_unit).length() > eps) return false;
        const vec3<R> d = (b.anchor + b.offset) - (a.anchor + a.offset);
        if ((d - unit * d.Dot(unit)).length() > eps * std::max<R>(R(1), d.length())) return false;  // in-plane shift
        pos.push_back(d.Dot(unit));
    }
    std::sort(pos.begin(), pos.end());
    for (size_t i = 1; i < pos.size(); ++i)
        if (std::abs((pos[i] - pos[i - 1]) - a.pxl_dz) > eps * std::max<R>(R(1), a.pxl_dz) * R(100)) return false;
    return true;
}
