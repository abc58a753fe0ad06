// MOCK: iDynTree::VectorDynSize, genuine members only
#pragma once
#include <vector>
namespace iDynTree {
class VectorDynSize {
    std::vector<double> m_data;
public:
    VectorDynSize() {}
    explicit VectorDynSize(unsigned int _size) : m_data(_size, 0.0) {}
    double operator()(const unsigned int index) const { return m_data[index]; }
    double& operator()(const unsigned int index) { return m_data[index]; }
    unsigned int size() const { return (unsigned int)m_data.size(); }
    const double* data() const { return m_data.data(); }
    double* data() { return m_data.data(); }
    void zero() { for (double& v : m_data) v = 0.0; }
    void resize(const unsigned int newSize) { m_data.assign(newSize, 0.0); }
};
} // namespace iDynTree
