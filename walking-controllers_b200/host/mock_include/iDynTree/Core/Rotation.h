// MOCK: iDynTree::Rotation (RotationRaw : MatrixFixSize<3, 3>, row-major), genuine members only
#pragma once
#include <cmath>
#include <iDynTree/Core/Position.h>
namespace iDynTree {
class Rotation {
    double m_data[9];
public:
    Rotation() { for (int i = 0; i < 9; i++) m_data[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    Rotation(double xx, double xy, double xz, double yx, double yy, double yz, double zx, double zy, double zz)
    { const double v[9] = {xx, xy, xz, yx, yy, yz, zx, zy, zz}; for (int i = 0; i < 9; i++) m_data[i] = v[i]; }
    double operator()(const unsigned int row, const unsigned int col) const { return m_data[3 * row + col]; }
    double& operator()(const unsigned int row, const unsigned int col) { return m_data[3 * row + col]; }
    unsigned int rows() const { return 3; }
    unsigned int cols() const { return 3; }
    const double* data() const { return m_data; }
    double* data() { return m_data; }
    Rotation inverse() const { Rotation t; for (unsigned r = 0; r < 3; r++) for (unsigned c = 0; c < 3; c++) t(r, c) = (*this)(c, r); return t; }
    Rotation operator*(const Rotation& o) const
    {
        Rotation t;
        for (unsigned r = 0; r < 3; r++) for (unsigned c = 0; c < 3; c++) { double s = 0; for (unsigned k = 0; k < 3; k++) s += (*this)(r, k) * o(k, c); t(r, c) = s; }
        return t;
    }
    Position operator*(const Position& p) const
    { return Position((*this)(0, 0) * p(0) + (*this)(0, 1) * p(1) + (*this)(0, 2) * p(2), (*this)(1, 0) * p(0) + (*this)(1, 1) * p(1) + (*this)(1, 2) * p(2),
                      (*this)(2, 0) * p(0) + (*this)(2, 1) * p(1) + (*this)(2, 2) * p(2)); }
    static Rotation RotZ(const double theta) { Rotation t; t(0, 0) = std::cos(theta); t(0, 1) = -std::sin(theta); t(1, 0) = std::sin(theta); t(1, 1) = std::cos(theta); return t; }
    static Rotation Identity() { return Rotation(); }
};
} // namespace iDynTree
