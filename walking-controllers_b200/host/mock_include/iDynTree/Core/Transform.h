// MOCK: iDynTree::Transform, genuine members only
#pragma once
#include <iDynTree/Core/Rotation.h>
namespace iDynTree {
class Transform {
    Position pos; Rotation rot;
public:
    Transform() {}
    Transform(const Rotation& _rot, const Position& origin) : pos(origin), rot(_rot) {}
    const Rotation& getRotation() const { return rot; }
    const Position& getPosition() const { return pos; }
    void setRotation(const Rotation& rotation) { rot = rotation; }
    void setPosition(const Position& position) { pos = position; }
    Transform inverse() const { const Rotation Ri = rot.inverse(); const Position p = Ri * pos; return Transform(Ri, Position(-p(0), -p(1), -p(2))); }
    Transform operator*(const Transform& o) const { return Transform(rot * o.rot, rot * o.pos + pos); }
    static Transform Identity() { return Transform(); }
};
} // namespace iDynTree
