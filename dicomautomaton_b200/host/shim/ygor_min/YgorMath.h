// Minimal stand-in for Ygor's YgorMath.h (github.com/hdclark/Ygor is not vendored by the reference and not installed
// here).  Only what the Demons path touches: vec3<T>.  Our own code, written from the usage sites in
// src/Alignment_Demons.{h,cc}, src/Alignment_Field.{h,cc} and their tests.  NOT a product component: the drop-in
// compiles against the real Ygor inside the reference tree; this shim only lets it compile and be tested here.
#pragma once
#include <cmath>
#include <iosfwd>
#include <list>
#include <map>
#include <string>
#include <ostream>
#include <istream>

template <class T>
class vec3 {
  public:
    T x = T(0), y = T(0), z = T(0);
    vec3() = default;
    vec3(T a, T b, T c) : x(a), y(b), z(c) {}
    vec3 operator+(const vec3 &o) const { return vec3(x + o.x, y + o.y, z + o.z); }
    vec3 operator-(const vec3 &o) const { return vec3(x - o.x, y - o.y, z - o.z); }
    vec3 operator*(T s) const { return vec3(x * s, y * s, z * s); }
    vec3 operator/(T s) const { return vec3(x / s, y / s, z / s); }
    vec3 &operator+=(const vec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3 &operator-=(const vec3 &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    bool operator==(const vec3 &o) const { return x == o.x && y == o.y && z == o.z; }
    T Dot(const vec3 &o) const { return x * o.x + y * o.y + z * o.z; }
    vec3 Cross(const vec3 &o) const { return vec3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
    T length() const { return std::sqrt(Dot(*this)); }
    T distance(const vec3 &o) const { return (*this - o).length(); }
    vec3 unit() const { const T l = length(); return (l > T(0)) ? (*this / l) : *this; }
    bool isfinite() const { return std::isfinite(x) && std::isfinite(y) && std::isfinite(z); }
};
template <class T>
std::ostream &operator<<(std::ostream &os, const vec3<T> &v) { return os << "(" << v.x << ", " << v.y << ", " << v.z << ")"; }

template <class T>
std::istream &operator>>(std::istream &is, vec3<T> &v) {  // "(x, y, z)"
    char c = 0;
    is >> c >> v.x >> c >> v.y >> c >> v.z >> c;
    return is;
}

// closed planar polygons with metadata ("ROIName", "NormalizedROIName"), as WarpImages selects them
template <class T>
class contour_of_points {
  public:
    std::list<vec3<T>> points;
    bool closed = true;
    std::map<std::string, std::string> metadata;
};
template <class T>
class contour_collection {
  public:
    std::list<contour_of_points<T>> contours;
};

template <class T>
class point_set {
  public:
    std::list<vec3<T>> points;
};
