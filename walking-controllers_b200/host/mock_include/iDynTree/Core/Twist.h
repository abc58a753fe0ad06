// MOCK: iDynTree::Twist (SpatialMotionVector : SpatialVector), genuine members only -- note: no data()
#pragma once
#include <iDynTree/Core/VectorFixSize.h>
namespace iDynTree {
class Twist {
    Vector3 linearVec3, angularVec3;
public:
    Twist() {}
    Vector3& getLinearVec3() { return linearVec3; }
    Vector3& getAngularVec3() { return angularVec3; }
    const Vector3& getLinearVec3() const { return linearVec3; }
    const Vector3& getAngularVec3() const { return angularVec3; }
    double& operator()(const unsigned int index) { return index < 3 ? linearVec3(index) : angularVec3(index - 3); }
    double operator()(const unsigned int index) const { return index < 3 ? linearVec3(index) : angularVec3(index - 3); }
    unsigned int size() const { return 6; }
    void zero() { linearVec3.zero(); angularVec3.zero(); }
    Vector6 asVector() const { Vector6 v; for (unsigned i = 0; i < 6; i++) v(i) = (*this)(i); return v; }
};
} // namespace iDynTree
