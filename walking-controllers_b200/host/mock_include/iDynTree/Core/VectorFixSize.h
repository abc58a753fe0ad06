// MOCK (see ../../README.md): iDynTree::VectorFixSize, genuine members only
#pragma once
#include <cstring>
namespace iDynTree {
template <unsigned int VecSize>
class VectorFixSize {
protected:
    double m_data[VecSize];
public:
    VectorFixSize() { zero(); }
    VectorFixSize(const double* in_data, const unsigned int in_size) { zero(); std::memcpy(m_data, in_data, sizeof(double) * (in_size < VecSize ? in_size : VecSize)); }
    double operator()(const unsigned int index) const { return m_data[index]; }
    double& operator()(const unsigned int index) { return m_data[index]; }
    double getVal(const unsigned int index) const { return m_data[index]; }
    bool setVal(const unsigned int index, const double new_el) { if (index >= VecSize) return false; m_data[index] = new_el; return true; }
    unsigned int size() const { return VecSize; }
    const double* data() const { return m_data; }
    double* data() { return m_data; }
    void zero() { for (unsigned int i = 0; i < VecSize; i++) m_data[i] = 0.0; }
};
typedef VectorFixSize<2> Vector2;
typedef VectorFixSize<3> Vector3;
typedef VectorFixSize<4> Vector4;
typedef VectorFixSize<6> Vector6;
} // namespace iDynTree
