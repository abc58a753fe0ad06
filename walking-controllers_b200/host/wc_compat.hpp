// wc_compat.hpp — the iDynTree / YARP value types that appear in the signatures of the reference's hot-path classes
// (WalkingController, WalkingQPIK, WalkingFK, ...), in two interchangeable forms:
//  * default: minimal stand-ins defined here (iDynTree and YARP are not available in this environment);
//  * -DWC_USE_REAL_IDYNTREE: the real headers (<iDynTree/Core/...>, <yarp/os/...>).
// The drop-in classes use (a) members that exist with the same signature in both (operator(), size(), data() of the vector /
// matrix / rotation types, getRotation() / getPosition(), zero(), check(), isNull()) and (b) the free helpers of namespace
// wc_compat at the end of this file, which have one implementation per form: configuration access (cfg_number / cfg_list /
// cfg_string over yarp::os::Searchable) and the row-major 3x4 [R|p] view of a transform the C ABI uses (toH12 / fromH12).
// `make check_real_api` compiles every class with -DWC_USE_REAL_IDYNTREE against mock_include/, a header set that exposes only
// genuine iDynTree 0.11 / YARP 3.1 members.
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
// linear (0..2) then angular (3..5), MIXED representation; like the real class it has operator() but NO data()
class Twist {
    std::array<double, 6> m_{};
public:
    double& operator()(int i) { return m_[i]; }
    double operator()(int i) const { return m_[i]; }
    int size() const { return 6; }
    void zero() { m_.fill(0.0); }
};

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
#else // WC_USE_REAL_IDYNTREE
#include <iDynTree/Core/MatrixDynSize.h>
#include <iDynTree/Core/Position.h>
#include <iDynTree/Core/Rotation.h>
#include <iDynTree/Core/Transform.h>
#include <iDynTree/Core/Twist.h>
#include <iDynTree/Core/VectorDynSize.h>
#include <iDynTree/Core/VectorFixSize.h>
#include <yarp/os/Bottle.h>
#include <yarp/os/Searchable.h>
#include <yarp/os/Value.h>
#endif // WC_USE_REAL_IDYNTREE

namespace wc_compat {
// ---- row-major 3x4 [R|p], the layout of include/wc_capi.h (same code for both forms: genuine members only)
inline void toH12(const iDynTree::Transform& T, double* H)
{
    const iDynTree::Rotation& R = T.getRotation(); const iDynTree::Position& p = T.getPosition();
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) H[4 * r + c] = R(r, c); H[4 * r + 3] = p(r); }
}
inline iDynTree::Transform fromH12(const double* H)
{
    iDynTree::Rotation R; iDynTree::Position p;
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R(r, c) = H[4 * r + c]; p(r) = H[4 * r + 3]; }
    return iDynTree::Transform(R, p);
}
inline void twist6(const iDynTree::Twist& t, double* out) { for (int k = 0; k < 6; k++) out[k] = t(k); }
// ---- configuration access
#ifndef WC_USE_REAL_IDYNTREE
inline double cfg_number(const yarp::os::Searchable& c, const std::string& key, double dflt) { return c.number(key, dflt); }
inline std::vector<double> cfg_list(const yarp::os::Searchable& c, const std::string& key) { return c.list(key); }
inline std::string cfg_string(const yarp::os::Searchable& c, const std::string& key, const std::string& dflt) { return c.string(key, dflt); }
#else
inline double cfg_number(const yarp::os::Searchable& c, const std::string& key, double dflt)
{
    if (!c.check(key)) return dflt;
    const yarp::os::Value& v = c.find(key);
    return v.isBool() ? (v.asBool() ? 1.0 : 0.0) : v.asDouble();
}
inline void cfg_flatten(const yarp::os::Value& v, std::vector<double>& out)
{
    if (v.isList()) { yarp::os::Bottle* b = v.asList(); for (int i = 0; b && i < (int)b->size(); i++) cfg_flatten(b->get(i), out); }
    else if (!v.isNull() && !v.isString()) out.push_back(v.asDouble());
}
// nested lists flattened in reading order ("((0,0,7500), (1,1,7500))" -> {0,0,7500,1,1,7500})
inline std::vector<double> cfg_list(const yarp::os::Searchable& c, const std::string& key)
{
    std::vector<double> out;
    if (c.check(key)) cfg_flatten(c.find(key), out);
    return out;
}
inline std::string cfg_string(const yarp::os::Searchable& c, const std::string& key, const std::string& dflt)
{ return c.check(key) ? std::string(c.find(key).asString()) : dflt; }
#endif
} // namespace wc_compat

#endif
