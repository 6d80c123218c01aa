// mock of PLPSLAM/type.h (Eigen typedefs) with the accessors / operators the adapter uses
#pragma once
#include <cmath>
#include <vector>
namespace PLPSLAM {
struct Vec3_t {
    double v[3];
    Vec3_t() : v{0, 0, 0} {}
    Vec3_t(double a, double b, double c) : v{a, b, c} {}
    double &operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double *data() { return v; }
    const double *data() const { return v; }
    Vec3_t operator-(const Vec3_t &o) const { return Vec3_t(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vec3_t operator-() const { return Vec3_t(-v[0], -v[1], -v[2]); }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
struct Vec2_t { double v[2]; double &operator()(int i) { return v[i]; } double operator()(int i) const { return v[i]; } };
struct Vec6_t { double v[6]; double &operator()(int i) { return v[i]; } double operator()(int i) const { return v[i]; } };
struct Mat33_t {
    double m[9];
    double &operator()(int r, int c) { return m[r * 3 + c]; }
    double operator()(int r, int c) const { return m[r * 3 + c]; }
    Mat33_t transpose() const { Mat33_t t; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t.m[c * 3 + r] = m[r * 3 + c]; return t; }
    Mat33_t operator-() const { Mat33_t t; for (int i = 0; i < 9; ++i) t.m[i] = -m[i]; return t; }
    Vec3_t operator*(const Vec3_t &x) const { return Vec3_t(m[0] * x(0) + m[1] * x(1) + m[2] * x(2), m[3] * x(0) + m[4] * x(1) + m[5] * x(2), m[6] * x(0) + m[7] * x(1) + m[8] * x(2)); }
};
template <int R, int C> struct block_result;
template <> struct block_result<3, 3> { typedef Mat33_t type; };
template <> struct block_result<3, 1> { typedef Vec3_t type; };
struct Mat44_t {
    double m[16];
    double &operator()(int r, int c) { return m[r * 4 + c]; }
    double operator()(int r, int c) const { return m[r * 4 + c]; }
    template <int R, int C> typename block_result<R, C>::type block(int, int) const { return typename block_result<R, C>::type(); }
};
template <class T> using eigen_alloc_vector = std::vector<T>;
}  // namespace PLPSLAM
