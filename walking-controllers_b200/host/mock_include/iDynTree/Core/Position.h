// MOCK: iDynTree::Position (PositionRaw : VectorFixSize<3>), genuine members only
#pragma once
#include <iDynTree/Core/VectorFixSize.h>
namespace iDynTree {
class Position : public VectorFixSize<3> {
public:
    Position() {}
    Position(double x, double y, double z) { m_data[0] = x; m_data[1] = y; m_data[2] = z; }
    Position operator+(const Position& o) const { return Position(m_data[0] + o(0), m_data[1] + o(1), m_data[2] + o(2)); }
    Position operator-(const Position& o) const { return Position(m_data[0] - o(0), m_data[1] - o(1), m_data[2] - o(2)); }
    static Position Zero() { return Position(); }
};
} // namespace iDynTree
