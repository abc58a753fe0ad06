// MOCK: iDynTree::MatrixDynSize (row-major storage), genuine members only
#pragma once
#include <vector>
namespace iDynTree {
class MatrixDynSize {
    std::vector<double> m_data; unsigned int m_rows = 0, m_cols = 0;
public:
    MatrixDynSize() {}
    MatrixDynSize(unsigned int _rows, unsigned int _cols) : m_data((size_t)_rows * _cols, 0.0), m_rows(_rows), m_cols(_cols) {}
    double operator()(const unsigned int row, const unsigned int col) const { return m_data[(size_t)row * m_cols + col]; }
    double& operator()(const unsigned int row, const unsigned int col) { return m_data[(size_t)row * m_cols + col]; }
    unsigned int rows() const { return m_rows; }
    unsigned int cols() const { return m_cols; }
    const double* data() const { return m_data.data(); }
    double* data() { return m_data.data(); }
    void zero() { for (double& v : m_data) v = 0.0; }
    void resize(const unsigned int _newRows, const unsigned int _newCols) { m_data.assign((size_t)_newRows * _newCols, 0.0); m_rows = _newRows; m_cols = _newCols; }
};
} // namespace iDynTree
