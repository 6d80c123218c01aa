// oracle/g2o_lite.hpp -- minimal restatement of the g2o pieces the reference's optimisers use
// (TEST INFRASTRUCTURE ONLY).
//
// g2o is a third-party dependency that is NOT under /root/reference and not installed in this environment
// (the reference finds it with an unversioned find_package(g2o), src/PLPSLAM/CMakeLists.txt:10-21; the README
// points at the author's own g2o fork).  PARITY UNPINNED: everything here is restated from g2o's published
// algorithm (SE3Quat, BaseUnary/BinaryEdge numeric Jacobians, RobustKernelHuber,
// OptimizationAlgorithmLevenberg, BlockSolver Schur complement) and anchored only on the reference's call
// sites (optimize/pose_optimizer.cc, optimize/local_bundle_adjuster*.cc, optimize/g2o/**).
#pragma once
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace g2o_lite {

struct Vec3 {
    double v[3];
    double &operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
};
inline Vec3 operator+(const Vec3 &a, const Vec3 &b) { return {{a[0] + b[0], a[1] + b[1], a[2] + b[2]}}; }
inline Vec3 operator-(const Vec3 &a, const Vec3 &b) { return {{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
inline Vec3 operator*(double s, const Vec3 &a) { return {{s * a[0], s * a[1], s * a[2]}}; }
inline double dot(const Vec3 &a, const Vec3 &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline Vec3 cross(const Vec3 &a, const Vec3 &b) {
    return {{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}};
}
inline double norm(const Vec3 &a) { return std::sqrt(dot(a, a)); }

struct Mat3 {
    double m[9];  // row-major
    double &operator()(int r, int c) { return m[r * 3 + c]; }
    double operator()(int r, int c) const { return m[r * 3 + c]; }
    static Mat3 identity() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
};
inline Mat3 operator*(const Mat3 &a, const Mat3 &b) {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
    return r;
}
inline Vec3 operator*(const Mat3 &a, const Vec3 &x) {
    return {{a(0, 0) * x[0] + a(0, 1) * x[1] + a(0, 2) * x[2], a(1, 0) * x[0] + a(1, 1) * x[1] + a(1, 2) * x[2],
             a(2, 0) * x[0] + a(2, 1) * x[1] + a(2, 2) * x[2]}};
}
inline Mat3 skew(const Vec3 &t) { return {{0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0}}; }
inline Mat3 transpose(const Mat3 &a) {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r(i, j) = a(j, i);
    return r;
}

// Eigen::Quaternion (w, x, y, z)
struct Quat {
    double w, x, y, z;
};
inline Quat quat_mul(const Quat &a, const Quat &b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline Quat quat_normalized(Quat q) {
    const double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return {q.w / n, q.x / n, q.y / n, q.z / n};
}
// Eigen::Quaternion(Matrix3) (Shepperd)
inline Quat quat_from_matrix(const Mat3 &m) {
    Quat q;
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (m(2, 1) - m(1, 2)) * t;
        q.y = (m(0, 2) - m(2, 0)) * t;
        q.z = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (m(k, j) - m(j, k)) * t;
        v[j] = (m(j, i) + m(i, j)) * t;
        v[k] = (m(k, i) + m(i, k)) * t;
        q.x = v[0];
        q.y = v[1];
        q.z = v[2];
    }
    return q;
}
inline Mat3 quat_to_matrix(const Quat &q) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    return {{1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx,
             1 - (txx + tyy)}};
}

// g2o::SE3Quat
struct SE3 {
    Quat q{1, 0, 0, 0};
    Vec3 t{{0, 0, 0}};
    void normalize_rotation() {  // SE3Quat::normalizeRotation
        if (q.w < 0) q = {-q.w, -q.x, -q.y, -q.z};
        q = quat_normalized(q);
    }
    Mat3 R() const { return quat_to_matrix(q); }
};
inline SE3 se3_from_matrix(const double *T /*4x4 row-major*/) {  // util/converter.cc:41-51 to_g2o_SE3
    Mat3 R;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R(r, c) = T[r * 4 + c];
    SE3 s;
    s.q = quat_from_matrix(R);
    s.t = {{T[3], T[7], T[11]}};
    s.normalize_rotation();
    return s;
}
inline void se3_to_matrix(const SE3 &s, double *T) {
    const Mat3 R = s.R();
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[r * 4 + c] = R(r, c);
        T[r * 4 + 3] = s.t[r];
    }
    T[12] = T[13] = T[14] = 0;
    T[15] = 1;
}
inline SE3 se3_mul(const SE3 &a, const SE3 &b) {  // SE3Quat::operator*
    SE3 r;
    r.q = quat_mul(a.q, b.q);
    r.t = a.t + a.R() * b.t;
    r.normalize_rotation();
    return r;
}
inline SE3 se3_exp(const double *u /*omega(3), upsilon(3)*/) {  // SE3Quat::exp
    const Vec3 omega{{u[0], u[1], u[2]}}, upsilon{{u[3], u[4], u[5]}};
    const double theta = norm(omega);
    const Mat3 Omega = skew(omega);
    const Mat3 Omega2 = Omega * Omega;
    Mat3 R, V;
    const Mat3 I = Mat3::identity();
    if (theta < 0.00001) {
        for (int i = 0; i < 9; ++i) {
            R.m[i] = I.m[i] + Omega.m[i] + 0.5 * Omega2.m[i];
            V.m[i] = I.m[i] + 0.5 * Omega.m[i] + (1. / 6.) * Omega2.m[i];
        }
    } else {
        const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta);
        const double c = (theta - std::sin(theta)) / (theta * theta * theta);
        for (int i = 0; i < 9; ++i) {
            R.m[i] = I.m[i] + a * Omega.m[i] + b * Omega2.m[i];
            V.m[i] = I.m[i] + b * Omega.m[i] + c * Omega2.m[i];
        }
    }
    SE3 s;
    s.q = quat_from_matrix(R);
    s.t = V * upsilon;
    s.normalize_rotation();
    return s;
}
// shot_vertex::oplusImpl (optimize/g2o/se3/shot_vertex.h:58-62)
inline SE3 se3_oplus(const SE3 &est, const double *u) { return se3_mul(se3_exp(u), est); }

// RobustKernelHuber::robustify
inline void huber(double e2, double delta, double rho[3]) {
    const double dsqr = delta * delta;
    if (e2 <= dsqr) {
        rho[0] = e2;
        rho[1] = 1.;
        rho[2] = 0.;
    } else {
        const double sqrte = std::sqrt(e2);
        rho[0] = 2 * sqrte * delta - dsqr;
        rho[1] = delta / sqrte;
        rho[2] = -0.5 * rho[1] / e2;
    }
}

struct Cam {
    double fx, fy, cx, cy, bf;
};

// ---- point edges (optimize/g2o/se3/perspective_pose_opt_edge.{h,cc}, perspective_reproj_edge.{h,cc})
// error = obs - cam_project(R X + t); dim 2 (mono) or 3 (stereo)
inline void point_error(const Cam &c, const Mat3 &R, const Vec3 &t, const Vec3 &Xw, const double *obs, bool stereo,
                        double *e, Vec3 *pc_out = nullptr) {
    const Vec3 pc = R * Xw + t;
    if (pc_out) *pc_out = pc;
    const double rx = c.fx * pc[0] / pc[2] + c.cx;
    e[0] = obs[0] - rx;
    e[1] = obs[1] - (c.fy * pc[1] / pc[2] + c.cy);
    if (stereo) e[2] = obs[2] - (rx - c.bf / pc[2]);
}
// d e / d pose (perspective_pose_opt_edge.cc:76-101, :142-173) -- rows x 6
inline void point_jac_pose(const Cam &c, const Vec3 &pc, bool stereo, double *J /*3x6 row-major*/) {
    const double x = pc[0], y = pc[1], z = pc[2], z_sq = z * z;
    J[0] = x * y / z_sq * c.fx;
    J[1] = -(1.0 + (x * x / z_sq)) * c.fx;
    J[2] = y / z * c.fx;
    J[3] = -1.0 / z * c.fx;
    J[4] = 0;
    J[5] = x / z_sq * c.fx;
    J[6] = (1.0 + y * y / z_sq) * c.fy;
    J[7] = -x * y / z_sq * c.fy;
    J[8] = -x / z * c.fy;
    J[9] = 0.0;
    J[10] = -1.0 / z * c.fy;
    J[11] = y / z_sq * c.fy;
    if (stereo) {
        J[12] = J[0] - c.bf * y / z_sq;
        J[13] = J[1] + c.bf * x / z_sq;
        J[14] = J[2];
        J[15] = J[3];
        J[16] = 0.0;
        J[17] = J[5] - c.bf / z_sq;
    }
}
// d e / d landmark (perspective_reproj_edge.cc:78-125, :166-214) -- rows x 3
inline void point_jac_landmark(const Cam &c, const Mat3 &R, const Vec3 &pc, bool stereo, double *J /*3x3*/) {
    const double x = pc[0], y = pc[1], z = pc[2], z_sq = z * z;
    for (int k = 0; k < 3; ++k) {
        J[k] = -c.fx * R(0, k) / z + c.fx * x * R(2, k) / z_sq;
        J[3 + k] = -c.fy * R(1, k) / z + c.fy * y * R(2, k) / z_sq;
        if (stereo) J[6 + k] = J[k] - c.bf * R(2, k) / z_sq;
    }
}

// ---- line edges (optimize/g2o/se3/pose_opt_edge_line3d_orthonormal.h:61-89,
//      reproj_edge_line3d_orthonormal.h:62-90)
inline Vec3 line_project(const Cam &c, const Mat3 &R, const Vec3 &t, const double *plucker) {
    const Vec3 n{{plucker[0], plucker[1], plucker[2]}}, d{{plucker[3], plucker[4], plucker[5]}};
    // (transformation_line_cw * L).head<3>() = R n + [t]x R d
    const Vec3 lc = R * n + (skew(t) * R) * d;
    // _K = [fy 0 0; 0 fx 0; -fy cx, -fx cy, fx fy]
    return {{c.fy * lc[0] + 0.0 * lc[1] + 0.0 * lc[2], 0.0 * lc[0] + c.fx * lc[1] + 0.0 * lc[2],
             -c.fy * c.cx * lc[0] + -c.fx * c.cy * lc[1] + c.fx * c.fy * lc[2]}};
}
inline void line_error(const Cam &c, const Mat3 &R, const Vec3 &t, const double *plucker, const double *obs /*xs,ys,xe,ye*/,
                       double *e) {
    const Vec3 p = line_project(c, R, t, plucker);
    e[0] = (obs[0] * p[0] + obs[1] * p[1] + p[2]) / std::sqrt(p[0] * p[0] + p[1] * p[1]);
    e[1] = (obs[2] * p[0] + obs[3] * p[1] + p[2]) / std::sqrt(p[0] * p[0] + p[1] * p[1]);
}

// dense Cholesky solve of an SPD system (LinearSolverEigen / CSparse are exact sparse Cholesky; any exact SPD
// solve is equivalent).  Returns false when the matrix is not positive definite.
inline bool cholesky_solve(std::vector<double> A /*n x n row-major, copied*/, const double *b, double *x, int n) {
    for (int j = 0; j < n; ++j) {
        double d = A[j * n + j];
        for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
        if (!(d > 0.0) || !std::isfinite(d)) return false;
        d = std::sqrt(d);
        A[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[i * n + j];
            for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = s / d;
        }
    }
    std::vector<double> y(n);
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= A[i * n + k] * y[k];
        y[i] = s / A[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < n; ++k) s -= A[k * n + i] * x[k];
        x[i] = s / A[i * n + i];
    }
    return true;
}

}  // namespace g2o_lite
