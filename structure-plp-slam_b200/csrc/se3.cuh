// se3.cuh -- SE(3) / quaternion / edge algebra shared by the pose optimiser and the bundle adjuster kernels.
//
// The reference delegates this to g2o (::g2o::SE3Quat, BaseVertex::oplus, RobustKernelHuber) and to its own
// edge classes under optimize/g2o/se3/.  Conventions reproduced here:
//   - pose vertex estimate = unit quaternion (w,x,y,z) + translation, T_cw;  update T <- exp(delta) * T with
//     delta = (omega, upsilon)                                  (optimize/g2o/se3/shot_vertex.h:58-62)
//   - SE3Quat::exp: Rodrigues for R, V matrix for the translation, small-angle branch at theta < 1e-5
//   - error = obs - projection                                   (perspective_pose_opt_edge.h:55-60)
#pragma once
#ifndef PLP_CTA_EMU
#include <cuda_runtime.h>
#endif
#include <math.h>

namespace plp {
namespace se3 {

struct Pose {  // unit quaternion + translation, plus the cached rotation matrix (row-major)
    double qw, qx, qy, qz;
    double t[3];
    double R[9];
};

__host__ __device__ inline void quat_to_R(double w, double x, double y, double z, double *R) {
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz);
    R[1] = txy - twz;
    R[2] = txz + twy;
    R[3] = txy + twz;
    R[4] = 1 - (txx + tzz);
    R[5] = tyz - twx;
    R[6] = txz - twy;
    R[7] = tyz + twx;
    R[8] = 1 - (txx + tyy);
}

__host__ __device__ inline void R_to_quat(const double *m, double &w, double &x, double &y, double &z) {
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        w = 0.5 * t;
        t = 0.5 / t;
        x = (m[7] - m[5]) * t;
        y = (m[2] - m[6]) * t;
        z = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 4]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[i * 4] - m[j * 4] - m[k * 4] + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        w = (m[k * 3 + j] - m[j * 3 + k]) * t;
        v[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        v[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
        x = v[0];
        y = v[1];
        z = v[2];
    }
}

__host__ __device__ inline void normalize(Pose &p) {  // SE3Quat::normalizeRotation + cache R
    if (p.qw < 0) {
        p.qw = -p.qw;
        p.qx = -p.qx;
        p.qy = -p.qy;
        p.qz = -p.qz;
    }
    const double n = sqrt(p.qw * p.qw + p.qx * p.qx + p.qy * p.qy + p.qz * p.qz);
    p.qw /= n;
    p.qx /= n;
    p.qy /= n;
    p.qz /= n;
    quat_to_R(p.qw, p.qx, p.qy, p.qz, p.R);
}

__host__ __device__ inline Pose from_matrix(const double *T /*4x4 row-major*/) {  // util/converter.cc:41-51
    Pose p;
    double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    R_to_quat(R, p.qw, p.qx, p.qy, p.qz);
    p.t[0] = T[3];
    p.t[1] = T[7];
    p.t[2] = T[11];
    normalize(p);
    return p;
}

__host__ __device__ inline void to_matrix(const Pose &p, double *T) {
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[r * 4 + c] = p.R[r * 3 + c];
        T[r * 4 + 3] = p.t[r];
    }
    T[12] = T[13] = T[14] = 0;
    T[15] = 1;
}

// T <- exp(u) * T   (shot_vertex::oplusImpl)
__host__ __device__ inline Pose oplus(const Pose &est, const double *u) {
    const double wx = u[0], wy = u[1], wz = u[2];
    const double theta = sqrt(wx * wx + wy * wy + wz * wz);
    // Omega = skew(omega), Omega2 = Omega*Omega
    const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double O2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
    double a, b, c;
    if (theta < 0.00001) {
        a = 1.0;
        b = 0.5;
        c = 1. / 6.;
    } else {
        a = sin(theta) / theta;
        b = (1 - cos(theta)) / (theta * theta);
        c = (theta - sin(theta)) / (theta * theta * theta);
    }
    double Rd[9], V[9];
    for (int i = 0; i < 9; ++i) {
        const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        Rd[i] = I + a * O[i] + b * O2[i];
        V[i] = I + b * O[i] + c * O2[i];
    }
    Pose d;
    R_to_quat(Rd, d.qw, d.qx, d.qy, d.qz);
    for (int r = 0; r < 3; ++r) d.t[r] = V[r * 3] * u[3] + V[r * 3 + 1] * u[4] + V[r * 3 + 2] * u[5];
    normalize(d);
    // SE3Quat::operator*: r = d.r * est.r ; t = d.t + d.r * est.t
    Pose o;
    o.qw = d.qw * est.qw - d.qx * est.qx - d.qy * est.qy - d.qz * est.qz;
    o.qx = d.qw * est.qx + d.qx * est.qw + d.qy * est.qz - d.qz * est.qy;
    o.qy = d.qw * est.qy + d.qy * est.qw + d.qz * est.qx - d.qx * est.qz;
    o.qz = d.qw * est.qz + d.qz * est.qw + d.qx * est.qy - d.qy * est.qx;
    for (int r = 0; r < 3; ++r)
        o.t[r] = d.t[r] + d.R[r * 3] * est.t[0] + d.R[r * 3 + 1] * est.t[1] + d.R[r * 3 + 2] * est.t[2];
    normalize(o);
    return o;
}

struct Cam {
    double fx, fy, cx, cy, bf;
};

__host__ __device__ inline void map_point(const double *R, const double *t, const double *X, double *pc) {
    pc[0] = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    pc[1] = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    pc[2] = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
}

// e = obs - cam_project(pc)  (perspective_pose_opt_edge.h:55-77 / :91-113)
__host__ __device__ inline void point_error(const Cam &c, const double *pc, const double *obs, bool stereo, double *e) {
    const double rx = c.fx * pc[0] / pc[2] + c.cx;
    e[0] = obs[0] - rx;
    e[1] = obs[1] - (c.fy * pc[1] / pc[2] + c.cy);
    e[2] = stereo ? obs[2] - (rx - c.bf / pc[2]) : 0.0;
}

// d e / d pose, rows x 6 (perspective_pose_opt_edge.cc:76-101, :142-173)
__host__ __device__ inline void point_jac_pose(const Cam &c, const double *pc, bool stereo, double *J) {
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
    } else {
        J[12] = J[13] = J[14] = J[15] = J[16] = J[17] = 0.0;
    }
}

// d e / d landmark, rows x 3 (perspective_reproj_edge.cc:78-125, :166-214)
__host__ __device__ inline void point_jac_landmark(const Cam &c, const double *R, const double *pc, bool stereo, double *J) {
    const double x = pc[0], y = pc[1], z = pc[2], z_sq = z * z;
    for (int k = 0; k < 3; ++k) {
        J[k] = -c.fx * R[k] / z + c.fx * x * R[6 + k] / z_sq;
        J[3 + k] = -c.fy * R[3 + k] / z + c.fy * y * R[6 + k] / z_sq;
        J[6 + k] = stereo ? J[k] - c.bf * R[6 + k] / z_sq : 0.0;
    }
}

// 2-D line l = K_l (R n + [t]x R d), K_l = [fy 0 0; 0 fx 0; -fy cx, -fx cy, fx fy]
// (pose_opt_edge_line3d_orthonormal.h:61-89, pose_opt_edge_wrapper.h:275-277)
__host__ __device__ inline void line_error(const Cam &c, const double *R, const double *t, const double *L /*n,d*/,
                                           const double *obs /*xs,ys,xe,ye*/, double *e) {
    double Rn[3], Rd[3];
    for (int r = 0; r < 3; ++r) {
        Rn[r] = R[r * 3] * L[0] + R[r * 3 + 1] * L[1] + R[r * 3 + 2] * L[2];
        Rd[r] = R[r * 3] * L[3] + R[r * 3 + 1] * L[4] + R[r * 3 + 2] * L[5];
    }
    const double lc0 = Rn[0] + (t[1] * Rd[2] - t[2] * Rd[1]);
    const double lc1 = Rn[1] + (t[2] * Rd[0] - t[0] * Rd[2]);
    const double lc2 = Rn[2] + (t[0] * Rd[1] - t[1] * Rd[0]);
    const double p0 = c.fy * lc0, p1 = c.fx * lc1, p2 = -c.fy * c.cx * lc0 - c.fx * c.cy * lc1 + c.fx * c.fy * lc2;
    const double den = sqrt(p0 * p0 + p1 * p1);
    e[0] = (obs[0] * p0 + obs[1] * p1 + p2) / den;
    e[1] = (obs[2] * p0 + obs[3] * p1 + p2) / den;
}

// RobustKernelHuber::robustify: rho[0] = rho(e2), rho[1] = rho'(e2)
__host__ __device__ inline void huber(double e2, double delta, double &rho0, double &rho1) {
    const double dsqr = delta * delta;
    if (e2 <= dsqr) {
        rho0 = e2;
        rho1 = 1.;
    } else {
        const double sqrte = sqrt(e2);
        rho0 = 2 * sqrte * delta - dsqr;
        rho1 = delta / sqrte;
    }
}

}  // namespace se3
}  // namespace plp
