// wc_compat.hpp — minimal stand-ins for the iDynTree / YARP value types that appear in the signatures
// of the reference's hot-path classes (WalkingController, WalkingQPIK, WalkingFK).  iDynTree and YARP are
// not available in this environment; with the real libraries present define WC_USE_REAL_IDYNTREE and
// include their headers instead — the drop-in classes only use the tiny API subset declared here
// (operator(), size(), data(), getRotation()/getPosition(), find()/check()).
#ifndef WC_COMPAT_HPP
#define WC_COMPAT_HPP

#include <array>
#include <cmath>
#include <cstddef>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#ifndef WC_USE_REAL_IDYNTREE
namespace iDynTree {

template <int N>
class VectorFixSize {
    std::array<double, N> m_{};
public:
    double& operator()(int i) { return m_[i]; }
    double operator()(int i) const { return m_[i]; }
    int size() const { return N; }
    double* data() { return m_.data(); }
    const double* data() const { return m_.data(); }
    void zero() { m_.fill(0.0); }
};
using Vector2 = VectorFixSize<2>;
using Vector3 = VectorFixSize<3>;
using Position = VectorFixSize<3>;
using Twist = VectorFixSize<6>; // linear (0..2) then angular (3..5), MIXED representation

class Rotation {
    std::array<double, 9> m_{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; // row-major
public:
    double& operator()(int r, int c) { return m_[3 * r + c]; }
    double operator()(int r, int c) const { return m_[3 * r + c]; }
    const double* data() const { return m_.data(); }
    double* data() { return m_.data(); }
    Rotation inverse() const { Rotation t; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) t(r, c) = (*this)(c, r); return t; }
    Rotation operator*(const Rotation& o) const
    {
        Rotation t;
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { double s = 0; for (int k = 0; k < 3; k++) s += (*this)(r, k) * o(k, c); t(r, c) = s; }
        return t;
    }
    static Rotation RotZ(double a) { Rotation t; t(0, 0) = std::cos(a); t(0, 1) = -std::sin(a); t(1, 0) = std::sin(a); t(1, 1) = std::cos(a); return t; }
};

class Transform {
    Rotation R_; Position p_;
public:
    Transform() { p_.zero(); }
    Transform(const Rotation& R, const Position& p) : R_(R), p_(p) {}
    const Rotation& getRotation() const { return R_; }
    const Position& getPosition() const { return p_; }
    void setRotation(const Rotation& R) { R_ = R; }
    void setPosition(const Position& p) { p_ = p; }
    // row-major 3x4 [R|p], the layout of include/wc_capi.h
    void toH12(double* H) const { for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) H[4 * r + c] = R_(r, c); H[4 * r + 3] = p_(r); } }
    static Transform fromH12(const double* H)
    {
        Transform t;
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) t.R_(r, c) = H[4 * r + c]; t.p_(r) = H[4 * r + 3]; }
        return t;
    }
};

class VectorDynSize {
    std::vector<double> m_;
public:
    VectorDynSize() = default;
    explicit VectorDynSize(size_t n) : m_(n, 0.0) {}
    void resize(size_t n) { m_.assign(n, 0.0); }
    size_t size() const { return m_.size(); }
    double& operator()(size_t i) { return m_[i]; }
    double operator()(size_t i) const { return m_[i]; }
    double* data() { return m_.data(); }
    const double* data() const { return m_.data(); }
};

class MatrixDynSize { // row-major, as iDynTree's
    std::vector<double> m_; size_t r_ = 0, c_ = 0;
public:
    MatrixDynSize() = default;
    MatrixDynSize(size_t r, size_t c) : m_(r * c, 0.0), r_(r), c_(c) {}
    void resize(size_t r, size_t c) { m_.assign(r * c, 0.0); r_ = r; c_ = c; }
    size_t rows() const { return r_; }
    size_t cols() const { return c_; }
    double& operator()(size_t r, size_t c) { return m_[r * c_ + c]; }
    double operator()(size_t r, size_t c) const { return m_[r * c_ + c]; }
    double* data() { return m_.data(); }
    const double* data() const { return m_.data(); }
};

} // namespace iDynTree

namespace yarp { namespace os {
// key -> flat list of numbers; nested lists in the .ini files are flattened in reading order
// ("((0,0,7500), (1,1,7500))" -> {0,0,7500,1,1,7500}), which is all the hot-path classes need.
class Searchable {
    std::map<std::string, std::vector<double>> num_;
    std::map<std::string, std::string> str_;
public:
    bool isNull() const { return num_.empty() && str_.empty(); }
    bool check(const std::string& k) const { return num_.count(k) || str_.count(k); }
    const std::vector<double>& list(const std::string& k) const { static const std::vector<double> e; auto it = num_.find(k); return it == num_.end() ? e : it->second; }
    double number(const std::string& k, double dflt) const { auto it = num_.find(k); return (it == num_.end() || it->second.empty()) ? dflt : it->second[0]; }
    std::string string(const std::string& k, const std::string& dflt) const { auto it = str_.find(k); return it == str_.end() ? dflt : it->second; }
    void put(const std::string& k, const std::vector<double>& v) { num_[k] = v; }
    void put(const std::string& k, double v) { num_[k] = {v}; }
    void put(const std::string& k, const std::string& v) { str_[k] = v; }
    // merges `key value...` lines of one reference .ini file (group headers / includes are ignored:
    // the reference passes each class its own group united with GENERAL, WalkingModule.cpp:192-193)
    bool fromIniFile(const std::string& path)
    {
        std::ifstream f(path);
        if (!f) return false;
        std::string line, pending_key; std::string pending;
        auto flush = [&]() {
            if (pending_key.empty()) return;
            std::string t = pending; bool numeric = true; std::vector<double> vals;
            for (char& ch : t) if (ch == '(' || ch == ')' || ch == ',') ch = ' ';
            std::istringstream is(t); std::string tok;
            while (is >> tok) { char* e = nullptr; double v = std::strtod(tok.c_str(), &e); if (*e) { numeric = false; break; } vals.push_back(v); }
            if (numeric && !vals.empty()) num_[pending_key] = vals; else str_[pending_key] = pending;
            pending_key.clear(); pending.clear();
        };
        int depth = 0;
        while (std::getline(f, line)) {
            const size_t hash = line.find('#'); if (hash != std::string::npos) line = line.substr(0, hash);
            if (line.find_first_not_of(" \t\r") == std::string::npos) continue;
            if (depth == 0) {
                flush();
                if (line[line.find_first_not_of(" \t")] == '[') continue;
                std::istringstream is(line); is >> pending_key; std::getline(is, pending);
            } else pending += " " + line;
            depth = 0; for (char ch : pending) { if (ch == '(') depth++; if (ch == ')') depth--; }
        }
        flush();
        return true;
    }
};
}} // namespace yarp::os
#endif // WC_USE_REAL_IDYNTREE

#endif
