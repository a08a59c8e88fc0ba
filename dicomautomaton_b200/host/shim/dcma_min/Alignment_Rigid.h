// Stand-in for the part of the reference's Alignment_Rigid.h / Ygor's affine_transform that WarpImages.cc touches
// (invert, apply_to on a point, coeff).  The Demons unit tests include the header but use nothing from it.  NOT a product
// component.
#pragma once
#include <array>
#include <stdexcept>

#include "YgorMath.h"

template <class T>
class affine_transform {
    std::array<std::array<T, 4>, 3> m_{{{T(1), T(0), T(0), T(0)}, {T(0), T(1), T(0), T(0)}, {T(0), T(0), T(1), T(0)}}};

  public:
    T &coeff(int r, int c) { return m_.at(static_cast<size_t>(r)).at(static_cast<size_t>(c)); }
    const T &coeff(int r, int c) const { return m_.at(static_cast<size_t>(r)).at(static_cast<size_t>(c)); }
    void apply_to(vec3<T> &v) const {
        const vec3<T> o(m_[0][0] * v.x + m_[0][1] * v.y + m_[0][2] * v.z + m_[0][3], m_[1][0] * v.x + m_[1][1] * v.y + m_[1][2] * v.z + m_[1][3],
                        m_[2][0] * v.x + m_[2][1] * v.y + m_[2][2] * v.z + m_[2][3]);
        v = o;
    }
    affine_transform invert() const {
        const auto &a = m_;
        const T det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                      a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
        if (det == T(0)) throw std::runtime_error("affine transform is singular");
        affine_transform o;
        o.m_[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) / det;
        o.m_[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) / det;
        o.m_[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) / det;
        o.m_[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) / det;
        o.m_[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) / det;
        o.m_[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) / det;
        o.m_[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) / det;
        o.m_[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) / det;
        o.m_[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) / det;
        for (int r = 0; r < 3; ++r) o.m_[r][3] = -(o.m_[r][0] * a[0][3] + o.m_[r][1] * a[1][3] + o.m_[r][2] * a[2][3]);
        return o;
    }
};
