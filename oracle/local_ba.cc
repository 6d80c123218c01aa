// oracle/local_ba.cc -- CPU restatement of optimize::local_bundle_adjuster, _extended_line and _extended_plane
// (TEST INFRASTRUCTURE ONLY).  Follows optimize/local_bundle_adjuster.cc:160-410,
// optimize/local_bundle_adjuster_extended_line.cc:190-640 and optimize/local_bundle_adjuster_extended_plane.cc
// :300-430 from the point where the graph has been gathered (the pointer-graph walk of :72-158 stays in the host
// adapter).  g2o's BlockSolver (Schur complement over the marginalised landmark vertices) and
// OptimizationAlgorithmLevenberg are restated in g2o_lite.hpp / here: PARITY UNPINNED (no g2o in this environment).
#include <algorithm>
#include <cstdint>

#include "g2o_lite.hpp"
#include "local_ba.h"

using namespace g2o_lite;

namespace {

// per-try trace of the most recent solve on this thread: (lambda used, rho, accepted) -- read by tests/test_ba_oracle.py
thread_local std::vector<double> g_lm_trace;
// step of g2o's central-difference Jacobians (BaseBinaryEdge::linearizeOplus: delta = 1e-9).  tests/ may move it by a relative
// 1e-7 to find the landmarks whose optimum the reference's own numeric differentiation does not determine reproducibly.
double g_numeric_delta = 1e-9;

// ---- optimize/g2o/line3d.h:57-207 -------------------------------------------------------------------------
struct Line3D {
    double v[6];  // (w = moment, d = direction)
    Vec3 w() const { return {{v[0], v[1], v[2]}}; }
    Vec3 d() const { return {{v[3], v[4], v[5]}}; }
    void normalize() {  // line3d.h:159-163
        const double n = 1.0 / norm(d());
        for (double &x : v) x *= n;
    }
};
struct Ortho {
    Mat3 U;
    double W[4];  // 2x2 row-major
};
Ortho to_orthonormal(const Line3D &line) {  // line3d.h:137-157
    Ortho o;
    const double mx = norm(line.d()), my = norm(line.w());
    const double wn = 1.0 / std::sqrt(mx * mx + my * my);
    o.W[0] = my * wn;
    o.W[1] = -mx * wn;
    o.W[2] = mx * wn;
    o.W[3] = my * wn;
    const double mn = 1.0 / my, dn = 1.0 / mx;
    const Vec3 mdcross = cross(line.w(), line.d());
    const double mdcrossn = 1.0 / norm(mdcross);
    for (int r = 0; r < 3; ++r) {
        o.U(r, 0) = line.w()[r] * mn;
        o.U(r, 1) = line.d()[r] * dn;
        o.U(r, 2) = mdcross[r] * mdcrossn;
    }
    return o;
}
Line3D from_orthonormal(const Ortho &o) {  // line3d.h:116-134
    Line3D l;
    for (int r = 0; r < 3; ++r) {
        l.v[r] = o.U(r, 0) * o.W[0];
        l.v[3 + r] = o.U(r, 1) * o.W[2];
    }
    l.normalize();
    return l;
}
Line3D line_oplus(const Line3D &line, const double *v) {  // line3d.h:171-186
    Ortho est = to_orthonormal(line);
    const double c = std::cos(v[3]), s = std::sin(v[3]);
    const double Wu[4] = {c, -s, s, c};
    Quat q{std::sqrt(1 - (v[0] * v[0] + v[1] * v[1] + v[2] * v[2])), v[0], v[1], v[2]};
    q = quat_normalized(q);
    const Mat3 Uu = quat_to_matrix(q);
    est.U = est.U * Uu;
    const double W0 = est.W[0] * Wu[0] + est.W[1] * Wu[2], W1 = est.W[0] * Wu[1] + est.W[1] * Wu[3];
    const double W2 = est.W[2] * Wu[0] + est.W[3] * Wu[2], W3 = est.W[2] * Wu[1] + est.W[3] * Wu[3];
    est.W[0] = W0;
    est.W[1] = W1;
    est.W[2] = W2;
    est.W[3] = W3;
    Line3D r = from_orthonormal(est);
    r.normalize();
    return r;
}

// reproj_edge_line3d::depth_is_positive_via_endpoints_trimming (reproj_edge_line3d_orthonormal.h:97-177)
bool line_depth_positive(const Cam &c, const SE3 &pose, const Line3D &line, const double *obs) {
    const Mat3 R = pose.R();
    const Vec3 t = pose.t;
    const Vec3 proj = line_project(c, R, t, line.v);
    const double l1 = proj[0], l2 = proj[1], l3 = proj[2];
    const double sp[2] = {obs[0], obs[1]}, ep[2] = {obs[2], obs[3]};
    const double x_sp = -(sp[1] - (l2 / l1) * sp[0] + (l3 / l2)) * ((l1 * l2) / (l1 * l1 + l2 * l2));
    const double y_sp = -(l1 / l2) * x_sp - (l3 / l2);
    const double x_ep = -(ep[1] - (l2 / l1) * ep[0] + (l3 / l2)) * ((l1 * l2) / (l1 * l1 + l2 * l2));
    const double y_ep = -(l1 / l2) * x_ep - (l3 / l2);
    const double y_0sp = sp[1] - (l2 / l1) * sp[0], y_0ep = ep[1] - (l2 / l1) * ep[0];
    // P = K [R | t]
    double P[12];
    const double K[9] = {c.fx, 0, c.cx, 0, c.fy, c.cy, 0, 0, 1};
    for (int r = 0; r < 3; ++r)
        for (int col = 0; col < 4; ++col) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += K[r * 3 + k] * (col < 3 ? R(k, col) : t[k]);
            P[r * 4 + col] = s;
        }
    auto plane_of = [&](double xc, double yc, double y0, double *pl) {
        const Vec3 a{{xc, yc, 1.0}}, b{{0.0, y0, 1.0}};
        const Vec3 lt = cross(a, b);
        for (int col = 0; col < 4; ++col) pl[col] = P[0 * 4 + col] * lt[0] + P[1 * 4 + col] * lt[1] + P[2 * 4 + col] * lt[2];
    };
    double pl_sp[4], pl_ep[4];
    plane_of(x_sp, y_sp, y_0sp, pl_sp);
    plane_of(x_ep, y_ep, y_0ep, pl_ep);
    // Pluecker matrix [[m]x d; -d^T 0]
    const Vec3 m = line.w(), d = line.d();
    const Mat3 Sm = skew(m);
    auto intersect = [&](const double *pl, double *X) {
        for (int r = 0; r < 3; ++r) X[r] = Sm(r, 0) * pl[0] + Sm(r, 1) * pl[1] + Sm(r, 2) * pl[2] + d[r] * pl[3];
        X[3] = -(d[0] * pl[0] + d[1] * pl[1] + d[2] * pl[2]);
    };
    double Xs[4], Xe[4];
    intersect(pl_sp, Xs);
    intersect(pl_ep, Xe);
    auto depth = [&](const double *X) {
        const Vec3 pw{{X[0] / X[3], X[1] / X[3], X[2] / X[3]}};
        return R(2, 0) * pw[0] + R(2, 1) * pw[1] + R(2, 2) * pw[2] + t[2] * 1.0;
    };
    return 0 < depth(Xs) && 0 < depth(Xe);
}

struct Solver {
    Cam cam;
    int n_kf = 0, n_pts = 0, n_lines = 0;
    std::vector<SE3> kf;
    std::vector<uint8_t> kf_fixed;
    std::vector<int> kf_hidx;  // index among the free poses or -1
    int n_free = 0;
    std::vector<Vec3> pts;
    std::vector<Line3D> lines;

    struct PtEdge {
        int kf, lm;
        double obs[3];
        bool stereo;
        double info;
        bool robust = true;
        int level = 0;
        double err[3] = {0, 0, 0};
    };
    struct LnEdge {
        int kf, lm;
        double obs[4];
        double info;
        bool robust = true;
        int level = 0;
        double err[2] = {0, 0};
    };
    struct PlEdge {
        int lm;
        double fn[4];
        double err = 0;
    };
    std::vector<PtEdge> pe;
    std::vector<LnEdge> le;
    std::vector<PlEdge> ple;
    double delta_pt = 0, delta_line = 0, delta_plane = 1.0;

    static double chi2_of(const double *e, int D, double info) {
        double s = 0;
        for (int i = 0; i < D; ++i) s += e[i] * (info * e[i]);
        return s;
    }
    void err_pt(PtEdge &e) const { point_error(cam, kf[e.kf].R(), kf[e.kf].t, pts[e.lm], e.obs, e.stereo, e.err); }
    void err_ln(LnEdge &e, const SE3 &pose, const Line3D &l) const { line_error(cam, pose.R(), pose.t, l.v, e.obs, e.err); }
    void err_pl(PlEdge &e, const Vec3 &X) const {
        const Vec3 n{{e.fn[0], e.fn[1], e.fn[2]}};
        e.err = (dot(X, n) + e.fn[3]) / norm(n);
    }
    void compute_active_errors() {
        for (auto &e : pe)
            if (e.level == 0) err_pt(e);
        for (auto &e : le)
            if (e.level == 0) err_ln(e, kf[e.kf], lines[e.lm]);
        for (auto &e : ple) err_pl(e, pts[e.lm]);
    }
    double active_robust_chi2() const {
        double chi = 0, rho[3];
        for (const auto &e : pe) {
            if (e.level) continue;
            const double c2 = chi2_of(e.err, e.stereo ? 3 : 2, e.info);
            if (e.robust) {
                huber(c2, delta_pt, rho);
                chi += rho[0];
            } else
                chi += c2;
        }
        for (const auto &e : le) {
            if (e.level) continue;
            const double c2 = chi2_of(e.err, 2, e.info);
            if (e.robust) {
                huber(c2, delta_line, rho);
                chi += rho[0];
            } else
                chi += c2;
        }
        for (const auto &e : ple) {
            huber(e.err * e.err, delta_plane, rho);
            chi += rho[0];
        }
        return chi;
    }

    // linear system in block form
    std::vector<double> Hpp, bp;        // n_free x 36 (diagonal blocks), n_free x 6
    std::vector<double> Hll_p, bl_p;    // points: 9 / 3
    std::vector<double> Hll_l, bl_l;    // lines: 16 / 4
    struct Hpl {
        int hidx, lm;
        double m[24];  // 6 x D
    };
    std::vector<Hpl> hpl_p, hpl_l;
    std::vector<uint8_t> act_p, act_l;  // landmark has at least one active edge

    void build_system() {
        Hpp.assign((size_t)n_free * 36, 0.0);
        bp.assign((size_t)n_free * 6, 0.0);
        Hll_p.assign((size_t)n_pts * 9, 0.0);
        bl_p.assign((size_t)n_pts * 3, 0.0);
        Hll_l.assign((size_t)n_lines * 16, 0.0);
        bl_l.assign((size_t)n_lines * 4, 0.0);
        hpl_p.clear();
        hpl_l.clear();
        act_p.assign(n_pts, 0);
        act_l.assign(n_lines, 0);
        double rho[3];
        for (auto &e : pe) {
            if (e.level) continue;
            act_p[e.lm] = 1;
            const int D = e.stereo ? 3 : 2;
            const Mat3 R = kf[e.kf].R();
            const Vec3 pc = R * pts[e.lm] + kf[e.kf].t;
            double Jp[18], Jl[9];
            point_jac_pose(cam, pc, e.stereo, Jp);
            point_jac_landmark(cam, R, pc, e.stereo, Jl);
            double w = e.info;
            if (e.robust) {
                huber(chi2_of(e.err, D, e.info), delta_pt, rho);
                w *= rho[1];
            }
            const int h = kf_hidx[e.kf];
            for (int r = 0; r < D; ++r) {
                const double we = -w * e.err[r];
                for (int a = 0; a < 3; ++a) {
                    bl_p[e.lm * 3 + a] += Jl[r * 3 + a] * we;
                    for (int c = 0; c < 3; ++c) Hll_p[e.lm * 9 + a * 3 + c] += Jl[r * 3 + a] * w * Jl[r * 3 + c];
                }
                if (h >= 0)
                    for (int a = 0; a < 6; ++a) {
                        bp[h * 6 + a] += Jp[r * 6 + a] * we;
                        for (int c = 0; c < 6; ++c) Hpp[h * 36 + a * 6 + c] += Jp[r * 6 + a] * w * Jp[r * 6 + c];
                    }
            }
            if (h >= 0) {
                Hpl b;
                b.hidx = h;
                b.lm = e.lm;
                for (int a = 0; a < 6; ++a)
                    for (int c = 0; c < 3; ++c) {
                        double s = 0;
                        for (int r = 0; r < D; ++r) s += Jp[r * 6 + a] * w * Jl[r * 3 + c];
                        b.m[a * 3 + c] = s;
                    }
                hpl_p.push_back(b);
            }
        }
        for (auto &e : le) {  // numeric Jacobians for both vertices (BaseBinaryEdge::linearizeOplus, delta 1e-9)
            if (e.level) continue;
            act_l[e.lm] = 1;
            const double delta = g_numeric_delta, scalar = 1 / (2 * delta);
            double Jp[12], Jl[8];
            const double before[2] = {e.err[0], e.err[1]};
            for (int d = 0; d < 6; ++d) {
                double add[6] = {0, 0, 0, 0, 0, 0};
                add[d] = delta;
                err_ln(e, se3_oplus(kf[e.kf], add), lines[e.lm]);
                const double e1[2] = {e.err[0], e.err[1]};
                add[d] = -delta;
                err_ln(e, se3_oplus(kf[e.kf], add), lines[e.lm]);
                Jp[0 * 6 + d] = scalar * (e1[0] - e.err[0]);
                Jp[1 * 6 + d] = scalar * (e1[1] - e.err[1]);
            }
            for (int d = 0; d < 4; ++d) {
                double add[4] = {0, 0, 0, 0};
                add[d] = delta;
                err_ln(e, kf[e.kf], line_oplus(lines[e.lm], add));
                const double e1[2] = {e.err[0], e.err[1]};
                add[d] = -delta;
                err_ln(e, kf[e.kf], line_oplus(lines[e.lm], add));
                Jl[0 * 4 + d] = scalar * (e1[0] - e.err[0]);
                Jl[1 * 4 + d] = scalar * (e1[1] - e.err[1]);
            }
            e.err[0] = before[0];
            e.err[1] = before[1];
            double w = e.info;
            if (e.robust) {
                huber(chi2_of(e.err, 2, e.info), delta_line, rho);
                w *= rho[1];
            }
            const int h = kf_hidx[e.kf];
            for (int r = 0; r < 2; ++r) {
                const double we = -w * e.err[r];
                for (int a = 0; a < 4; ++a) {
                    bl_l[e.lm * 4 + a] += Jl[r * 4 + a] * we;
                    for (int c = 0; c < 4; ++c) Hll_l[e.lm * 16 + a * 4 + c] += Jl[r * 4 + a] * w * Jl[r * 4 + c];
                }
                if (h >= 0)
                    for (int a = 0; a < 6; ++a) {
                        bp[h * 6 + a] += Jp[r * 6 + a] * we;
                        for (int c = 0; c < 6; ++c) Hpp[h * 36 + a * 6 + c] += Jp[r * 6 + a] * w * Jp[r * 6 + c];
                    }
            }
            if (h >= 0) {
                Hpl b;
                b.hidx = h;
                b.lm = e.lm;
                for (int a = 0; a < 6; ++a)
                    for (int c = 0; c < 4; ++c) b.m[a * 4 + c] = Jp[0 * 6 + a] * w * Jl[0 * 4 + c] + Jp[1 * 6 + a] * w * Jl[1 * 4 + c];
                hpl_l.push_back(b);
            }
        }
        for (auto &e : ple) {  // unary on the landmark, numeric Jacobian (additive landmark update)
            act_p[e.lm] = 1;
            const double delta = g_numeric_delta, scalar = 1 / (2 * delta);
            double J[3];
            const double before = e.err;
            for (int d = 0; d < 3; ++d) {
                Vec3 X = pts[e.lm];
                X[d] += delta;
                err_pl(e, X);
                const double e1 = e.err;
                X = pts[e.lm];
                X[d] -= delta;
                err_pl(e, X);
                J[d] = scalar * (e1 - e.err);
            }
            e.err = before;
            huber(e.err * e.err, delta_plane, rho);
            const double w = rho[1];
            for (int a = 0; a < 3; ++a) {
                bl_p[e.lm * 3 + a] += J[a] * (-w * e.err);
                for (int c = 0; c < 3; ++c) Hll_p[e.lm * 9 + a * 3 + c] += J[a] * w * J[c];
            }
        }
    }

    static bool inv_spd(const double *A, int n, double *Ainv) {  // Gauss-Jordan on a small dense block
        double M[16], I[16];
        for (int i = 0; i < n * n; ++i) {
            M[i] = A[i];
            I[i] = 0;
        }
        for (int i = 0; i < n; ++i) I[i * n + i] = 1;
        for (int c = 0; c < n; ++c) {
            int piv = c;
            for (int r = c + 1; r < n; ++r)
                if (std::fabs(M[r * n + c]) > std::fabs(M[piv * n + c])) piv = r;
            if (M[piv * n + c] == 0.0) return false;
            if (piv != c)
                for (int k = 0; k < n; ++k) {
                    std::swap(M[c * n + k], M[piv * n + k]);
                    std::swap(I[c * n + k], I[piv * n + k]);
                }
            const double d = 1.0 / M[c * n + c];
            for (int k = 0; k < n; ++k) {
                M[c * n + k] *= d;
                I[c * n + k] *= d;
            }
            for (int r = 0; r < n; ++r) {
                if (r == c) continue;
                const double f = M[r * n + c];
                for (int k = 0; k < n; ++k) {
                    M[r * n + k] -= f * M[c * n + k];
                    I[r * n + k] -= f * I[c * n + k];
                }
            }
        }
        for (int i = 0; i < n * n; ++i) Ainv[i] = I[i];
        return true;
    }

    // BlockSolver::solve with Schur complement; x = (dp, dl_points, dl_lines)
    bool solve(double lambda, std::vector<double> &xp, std::vector<double> &xl_p, std::vector<double> &xl_l) {
        const int n = 6 * n_free;
        std::vector<double> S((size_t)n * n, 0.0), g(n, 0.0);
        for (int h = 0; h < n_free; ++h)
            for (int a = 0; a < 6; ++a) {
                g[h * 6 + a] = bp[h * 6 + a];
                for (int c = 0; c < 6; ++c) S[(size_t)(h * 6 + a) * n + h * 6 + c] = Hpp[h * 36 + a * 6 + c] + (a == c ? lambda : 0.0);
            }
        std::vector<double> Dinv_p((size_t)n_pts * 9), Dinv_l((size_t)n_lines * 16);
        auto schur = [&](int D, const std::vector<Hpl> &hpl, const std::vector<double> &Hll, const std::vector<double> &bl,
                         std::vector<double> &Dinv, const std::vector<uint8_t> &act, int n_lm) {
            // group the Hpl blocks by landmark
            std::vector<std::vector<int>> by_lm(n_lm);
            for (int i = 0; i < (int)hpl.size(); ++i) by_lm[hpl[i].lm].push_back(i);
            for (int l = 0; l < n_lm; ++l) {
                if (!act[l]) continue;
                double Dm[16];
                for (int i = 0; i < D * D; ++i) Dm[i] = Hll[(size_t)l * D * D + i];
                for (int i = 0; i < D; ++i) Dm[i * D + i] += lambda;
                if (!inv_spd(Dm, D, &Dinv[(size_t)l * D * D])) return false;
                const double *Di = &Dinv[(size_t)l * D * D];
                double db[4];
                for (int a = 0; a < D; ++a) {
                    db[a] = 0;
                    for (int c = 0; c < D; ++c) db[a] += Di[a * D + c] * bl[(size_t)l * D + c];
                }
                for (int i1 : by_lm[l]) {
                    const Hpl &B1 = hpl[i1];
                    double BD[24];
                    for (int a = 0; a < 6; ++a)
                        for (int c = 0; c < D; ++c) {
                            double s = 0;
                            for (int k = 0; k < D; ++k) s += B1.m[a * D + k] * Di[k * D + c];
                            BD[a * D + c] = s;
                        }
                    for (int a = 0; a < 6; ++a) {
                        double s = 0;
                        for (int c = 0; c < D; ++c) s += B1.m[a * D + c] * db[c];
                        g[B1.hidx * 6 + a] -= s;
                    }
                    for (int i2 : by_lm[l]) {
                        const Hpl &B2 = hpl[i2];
                        for (int a = 0; a < 6; ++a)
                            for (int c = 0; c < 6; ++c) {
                                double s = 0;
                                for (int k = 0; k < D; ++k) s += BD[a * D + k] * B2.m[c * D + k];
                                S[(size_t)(B1.hidx * 6 + a) * n + B2.hidx * 6 + c] -= s;
                            }
                    }
                }
            }
            return true;
        };
        if (!schur(3, hpl_p, Hll_p, bl_p, Dinv_p, act_p, n_pts)) return false;
        if (!schur(4, hpl_l, Hll_l, bl_l, Dinv_l, act_l, n_lines)) return false;
        xp.assign(n, 0.0);
        if (n > 0 && !cholesky_solve(S, g.data(), xp.data(), n)) return false;
        auto backsub = [&](int D, const std::vector<Hpl> &hpl, const std::vector<double> &bl, const std::vector<double> &Dinv,
                           const std::vector<uint8_t> &act, int n_lm, std::vector<double> &xl) {
            std::vector<double> cl(bl);
            for (const Hpl &B : hpl)
                for (int c = 0; c < D; ++c) {
                    double s = 0;
                    for (int a = 0; a < 6; ++a) s += B.m[a * D + c] * xp[B.hidx * 6 + a];
                    cl[(size_t)B.lm * D + c] -= s;
                }
            xl.assign((size_t)n_lm * D, 0.0);
            for (int l = 0; l < n_lm; ++l) {
                if (!act[l]) continue;
                for (int a = 0; a < D; ++a) {
                    double s = 0;
                    for (int c = 0; c < D; ++c) s += Dinv[(size_t)l * D * D + a * D + c] * cl[(size_t)l * D + c];
                    xl[(size_t)l * D + a] = s;
                }
            }
        };
        backsub(3, hpl_p, bl_p, Dinv_p, act_p, n_pts, xl_p);
        backsub(4, hpl_l, bl_l, Dinv_l, act_l, n_lines, xl_l);
        return true;
    }

    int optimize(int iterations, const volatile uint8_t *force_stop, int *lm_tries) {
        double lambda = 0, ni = 2;
        int done = 0;
        for (int it = 0; it < iterations; ++it) {
            if (force_stop && *force_stop) break;
            compute_active_errors();
            double current_chi = active_robust_chi2();
            double temp_chi = current_chi;
            build_system();
            if (it == 0) {  // computeLambdaInit over all (non-fixed, active) vertices
                double md = 0;
                for (int h = 0; h < n_free; ++h)
                    for (int j = 0; j < 6; ++j) md = std::max(std::fabs(Hpp[h * 36 + j * 6 + j]), md);
                for (int l = 0; l < n_pts; ++l)
                    if (act_p[l])
                        for (int j = 0; j < 3; ++j) md = std::max(std::fabs(Hll_p[l * 9 + j * 3 + j]), md);
                for (int l = 0; l < n_lines; ++l)
                    if (act_l[l])
                        for (int j = 0; j < 4; ++j) md = std::max(std::fabs(Hll_l[l * 16 + j * 4 + j]), md);
                lambda = 1e-5 * md;
                ni = 2;
            }
            double rho = 0;
            int qmax = 0;
            bool lambda_finite = true;
            do {
                const auto kf_backup = kf;
                const auto pts_backup = pts;
                const auto lines_backup = lines;
                std::vector<double> xp, xl_p, xl_l;
                const bool ok2 = solve(lambda, xp, xl_p, xl_l);
                if (ok2) {
                    for (int k = 0; k < n_kf; ++k)
                        if (kf_hidx[k] >= 0) kf[k] = se3_oplus(kf[k], &xp[kf_hidx[k] * 6]);
                    for (int l = 0; l < n_pts; ++l)
                        if (act_p[l])
                            for (int a = 0; a < 3; ++a) pts[l][a] += xl_p[l * 3 + a];
                    for (int l = 0; l < n_lines; ++l)
                        if (act_l[l]) lines[l] = line_oplus(lines[l], &xl_l[l * 4]);
                }
                compute_active_errors();
                temp_chi = active_robust_chi2();
                if (!ok2) temp_chi = std::numeric_limits<double>::max();
                rho = current_chi - temp_chi;
                double scale = 0;
                if (ok2) {
                    for (int j = 0; j < 6 * n_free; ++j) scale += xp[j] * (lambda * xp[j] + bp[j]);
                    for (int l = 0; l < n_pts; ++l)
                        if (act_p[l])
                            for (int a = 0; a < 3; ++a) scale += xl_p[l * 3 + a] * (lambda * xl_p[l * 3 + a] + bl_p[l * 3 + a]);
                    for (int l = 0; l < n_lines; ++l)
                        if (act_l[l])
                            for (int a = 0; a < 4; ++a) scale += xl_l[l * 4 + a] * (lambda * xl_l[l * 4 + a] + bl_l[l * 4 + a]);
                }
                scale += 1e-3;
                rho /= scale;
                if (lm_tries) ++*lm_tries;
                g_lm_trace.push_back(lambda);
                g_lm_trace.push_back(rho);
                g_lm_trace.push_back((rho > 0 && std::isfinite(temp_chi)) ? 1.0 : 0.0);
                if (rho > 0 && std::isfinite(temp_chi)) {
                    double alpha = 1. - std::pow((2 * rho - 1), 3);
                    alpha = std::min(alpha, 2. / 3.);
                    lambda *= std::max(1. / 3., alpha);
                    ni = 2;
                    current_chi = temp_chi;
                } else {
                    lambda *= ni;
                    ni *= 2;
                    kf = kf_backup;
                    pts = pts_backup;
                    lines = lines_backup;
                    if (!std::isfinite(lambda)) {
                        lambda_finite = false;
                        break;
                    }
                }
                qmax++;
            } while (rho < 0 && qmax < 10 && !(force_stop && *force_stop));
            ++done;
            if (qmax == 10 || rho == 0 || !lambda_finite) break;
        }
        return done;
    }
};

}  // namespace

extern "C" void orc_debug_set_numeric_delta(double d) { g_numeric_delta = d; }

extern "C" int orc_debug_lm_trace(double *out, int cap) {
    const int n = (int)g_lm_trace.size() / 3;
    for (int i = 0; i < 3 * std::min(n, cap); ++i) out[i] = g_lm_trace[i];
    return n;
}

extern "C" int orc_local_ba(const orc_ba_problem *p, int num_first_iter, int num_second_iter,
                            const volatile uint8_t *force_stop, orc_ba_result *r) {
    const float chi_sq_2D = 5.99146f, chi_sq_3D = 7.81473f;
    g_lm_trace.clear();
    Solver S;
    S.cam = {p->fx, p->fy, p->cx, p->cy, p->focal_x_baseline};
    S.n_kf = p->n_kf;
    S.n_pts = p->n_pts;
    S.n_lines = p->n_lines;
    S.kf.resize(p->n_kf);
    S.kf_fixed.assign(p->kf_fixed, p->kf_fixed + p->n_kf);
    S.kf_hidx.assign(p->n_kf, -1);
    for (int k = 0; k < p->n_kf; ++k) {
        S.kf[k] = se3_from_matrix(p->kf_pose_cw + 16 * (size_t)k);
        if (!p->kf_fixed[k]) S.kf_hidx[k] = S.n_free++;
    }
    S.pts.resize(p->n_pts);
    for (int l = 0; l < p->n_pts; ++l) S.pts[l] = {{p->pt_pos_w[3 * l], p->pt_pos_w[3 * l + 1], p->pt_pos_w[3 * l + 2]}};
    S.lines.resize(p->n_lines);
    for (int l = 0; l < p->n_lines; ++l)
        for (int k = 0; k < 6; ++k) S.lines[l].v[k] = p->line_plucker[6 * (size_t)l + k];
    S.delta_pt = p->setup_type == 0 ? std::sqrt(chi_sq_2D) : std::sqrt(chi_sq_3D);
    S.delta_line = std::sqrt(chi_sq_2D);
    S.pe.resize(p->n_pt_edges);
    for (int i = 0; i < p->n_pt_edges; ++i) {
        auto &e = S.pe[i];
        e.kf = p->pt_edge_kf[i];
        e.lm = p->pt_edge_lm[i];
        e.obs[0] = p->pt_edge_obs[3 * i];
        e.obs[1] = p->pt_edge_obs[3 * i + 1];
        e.obs[2] = p->pt_edge_obs[3 * i + 2];
        e.stereo = !(p->pt_edge_obs[3 * i + 2] < 0);
        e.info = p->pt_edge_inv_sigma_sq[i];
    }
    S.le.resize(p->n_line_edges);
    for (int i = 0; i < p->n_line_edges; ++i) {
        auto &e = S.le[i];
        e.kf = p->line_edge_kf[i];
        e.lm = p->line_edge_lm[i];
        for (int k = 0; k < 4; ++k) e.obs[k] = p->line_edge_obs[4 * i + k];
        e.info = p->line_edge_inv_sigma_sq[i];
    }
    S.ple.resize(p->n_plane_edges);
    for (int i = 0; i < p->n_plane_edges; ++i) {
        S.ple[i].lm = p->plane_edge_lm[i];
        for (int k = 0; k < 4; ++k) S.ple[i].fn[k] = p->plane_edge_fn[4 * i + k];
    }
    r->iters_first = r->iters_second = 0;
    r->lm_tries = 0;
    auto write_back = [&]() {
        for (int k = 0; k < p->n_kf; ++k) se3_to_matrix(S.kf[k], r->kf_pose_cw + 16 * (size_t)k);
        for (int l = 0; l < p->n_pts; ++l)
            for (int a = 0; a < 3; ++a) r->pt_pos_w[3 * l + a] = S.pts[l][a];
        for (int l = 0; l < p->n_lines; ++l)
            for (int k = 0; k < 6; ++k) r->line_plucker[6 * (size_t)l + k] = S.lines[l].v[k];
    };
    for (int i = 0; i < p->n_pt_edges; ++i) r->pt_edge_outlier[i] = 0;
    for (int i = 0; i < p->n_line_edges; ++i) r->line_edge_outlier[i] = 0;
    if (force_stop && *force_stop) {  // local_bundle_adjuster.cc:276-282: return before anything is written
        write_back();
        return 0;
    }
    r->iters_first = S.optimize(num_first_iter, force_stop, &r->lm_tries);
    const bool run_robust = !(force_stop && *force_stop);
    auto pt_is_outlier = [&](const Solver::PtEdge &e) {
        const double thr = e.stereo ? (double)chi_sq_3D : (double)chi_sq_2D;
        const Vec3 pc = S.kf[e.kf].R() * S.pts[e.lm] + S.kf[e.kf].t;
        return thr < Solver::chi2_of(e.err, e.stereo ? 3 : 2, e.info) || !(0.0 < pc[2]);
    };
    auto ln_is_outlier = [&](const Solver::LnEdge &e) {
        return (double)chi_sq_2D < Solver::chi2_of(e.err, 2, e.info) ||
               !line_depth_positive(S.cam, S.kf[e.kf], S.lines[e.lm], e.obs);
    };
    if (run_robust) {
        for (auto &e : S.pe) {  // local_bundle_adjuster.cc:303-332
            if (pt_is_outlier(e)) e.level = 1;
            e.robust = false;
        }
        for (auto &e : S.le) {
            if (ln_is_outlier(e)) e.level = 1;
            e.robust = false;
        }
        r->iters_second = S.optimize(num_second_iter, force_stop, &r->lm_tries);
    }
    for (int i = 0; i < p->n_pt_edges; ++i) r->pt_edge_outlier[i] = pt_is_outlier(S.pe[i]) ? 1 : 0;  // :342-372
    for (int i = 0; i < p->n_line_edges; ++i) r->line_edge_outlier[i] = ln_is_outlier(S.le[i]) ? 1 : 0;
    write_back();
    S.compute_active_errors();
    r->final_chi2 = S.active_robust_chi2();
    return r->iters_first + r->iters_second;
}

extern "C" int orc_global_ba(const orc_ba_problem *p, int num_iter, int use_huber_kernel,
                            const volatile uint8_t *force_stop, orc_ba_result *r) {
    const float chi_sq_2D = 5.99146f, chi_sq_3D = 7.81473f;
    g_lm_trace.clear();
    Solver S;
    S.cam = {p->fx, p->fy, p->cx, p->cy, p->focal_x_baseline};
    S.n_kf = p->n_kf;
    S.n_pts = p->n_pts;
    S.n_lines = p->n_lines;
    S.kf.resize(p->n_kf);
    S.kf_fixed.assign(p->kf_fixed, p->kf_fixed + p->n_kf);
    S.kf_hidx.assign(p->n_kf, -1);
    for (int k = 0; k < p->n_kf; ++k) {
        S.kf[k] = se3_from_matrix(p->kf_pose_cw + 16 * (size_t)k);
        if (!p->kf_fixed[k]) S.kf_hidx[k] = S.n_free++;
    }
    S.pts.resize(p->n_pts);
    for (int l = 0; l < p->n_pts; ++l) S.pts[l] = {{p->pt_pos_w[3 * l], p->pt_pos_w[3 * l + 1], p->pt_pos_w[3 * l + 2]}};
    S.lines.resize(p->n_lines);
    for (int l = 0; l < p->n_lines; ++l)
        for (int k = 0; k < 6; ++k) S.lines[l].v[k] = p->line_plucker[6 * (size_t)l + k];
    S.delta_pt = p->setup_type == 0 ? std::sqrt(chi_sq_2D) : std::sqrt(chi_sq_3D);
    S.delta_line = std::sqrt(chi_sq_2D);
    S.pe.resize(p->n_pt_edges);
    for (int i = 0; i < p->n_pt_edges; ++i) {
        auto &e = S.pe[i];
        e.kf = p->pt_edge_kf[i];
        e.lm = p->pt_edge_lm[i];
        e.obs[0] = p->pt_edge_obs[3 * i];
        e.obs[1] = p->pt_edge_obs[3 * i + 1];
        e.obs[2] = p->pt_edge_obs[3 * i + 2];
        e.stereo = !(p->pt_edge_obs[3 * i + 2] < 0);
        e.info = p->pt_edge_inv_sigma_sq[i];
    }
    S.le.resize(p->n_line_edges);
    for (int i = 0; i < p->n_line_edges; ++i) {
        auto &e = S.le[i];
        e.kf = p->line_edge_kf[i];
        e.lm = p->line_edge_lm[i];
        for (int k = 0; k < 4; ++k) e.obs[k] = p->line_edge_obs[4 * i + k];
        e.info = p->line_edge_inv_sigma_sq[i];
    }
    S.ple.resize(p->n_plane_edges);
    for (int i = 0; i < p->n_plane_edges; ++i) {
        S.ple[i].lm = p->plane_edge_lm[i];
        for (int k = 0; k < 4; ++k) S.ple[i].fn[k] = p->plane_edge_fn[4 * i + k];
    }
    // optimize/global_bundle_adjuster.cc:64-253: the same vertices / edges as the local adjusters over ALL keyframes and
    // landmarks (only keyframe id 0 fixed), one optimize(num_iter) with or without the Huber kernel, no outlier rounds
    r->iters_first = r->iters_second = 0;
    r->lm_tries = 0;
    for (int i = 0; i < p->n_pt_edges; ++i) r->pt_edge_outlier[i] = 0;
    for (int i = 0; i < p->n_line_edges; ++i) r->line_edge_outlier[i] = 0;
    if (!use_huber_kernel) {
        for (auto &e : S.pe) e.robust = false;
        for (auto &e : S.le) e.robust = false;
    }
    r->iters_first = S.optimize(num_iter, force_stop, &r->lm_tries);
    for (int k = 0; k < p->n_kf; ++k) se3_to_matrix(S.kf[k], r->kf_pose_cw + 16 * (size_t)k);
    for (int l = 0; l < p->n_pts; ++l)
        for (int a = 0; a < 3; ++a) r->pt_pos_w[3 * l + a] = S.pts[l][a];
    for (int l = 0; l < p->n_lines; ++l)
        for (int k = 0; k < 6; ++k) r->line_plucker[6 * (size_t)l + k] = S.lines[l].v[k];
    S.compute_active_errors();
    r->final_chi2 = S.active_robust_chi2();
    return r->iters_first;
}

// ---- parity taps for tests/test_ba_oracle.py: the edge algebra this file and pose_opt.cc build on --------------------
extern "C" void orc_debug_point_edge(const double *cam5 /*fx fy cx cy bf*/, const double *T_cw, const double *X,
                                     const double *obs3, double *e3, double *Jpose18, double *Jlm9) {
    const Cam c{cam5[0], cam5[1], cam5[2], cam5[3], cam5[4]};
    const SE3 P = se3_from_matrix(T_cw);
    const bool stereo = !(obs3[2] < 0);
    Vec3 pc;
    e3[2] = 0;
    point_error(c, P.R(), P.t, Vec3{{X[0], X[1], X[2]}}, obs3, stereo, e3, &pc);
    for (int i = 0; i < 18; ++i) Jpose18[i] = 0;
    for (int i = 0; i < 9; ++i) Jlm9[i] = 0;
    point_jac_pose(c, pc, stereo, Jpose18);
    point_jac_landmark(c, P.R(), pc, stereo, Jlm9);
}
extern "C" void orc_debug_se3_oplus(const double *T_cw, const double *u6, double *T_out) {
    se3_to_matrix(se3_oplus(se3_from_matrix(T_cw), u6), T_out);
}
extern "C" void orc_debug_line_oplus(const double *L6, const double *v4, double *out6) {
    Line3D l;
    for (int k = 0; k < 6; ++k) l.v[k] = L6[k];
    const Line3D r = line_oplus(l, v4);
    for (int k = 0; k < 6; ++k) out6[k] = r.v[k];
}
extern "C" void orc_debug_line_error(const double *cam5, const double *T_cw, const double *L6, const double *obs4, double *e2) {
    const Cam c{cam5[0], cam5[1], cam5[2], cam5[3], cam5[4]};
    const SE3 P = se3_from_matrix(T_cw);
    line_error(c, P.R(), P.t, L6, obs4, e2);
}
extern "C" int orc_debug_line_depth_positive(const double *cam5, const double *T_cw, const double *L6, const double *obs4) {
    const Cam c{cam5[0], cam5[1], cam5[2], cam5[3], cam5[4]};
    Line3D l;
    for (int k = 0; k < 6; ++k) l.v[k] = L6[k];
    return line_depth_positive(c, se3_from_matrix(T_cw), l, obs4) ? 1 : 0;
}

// ---- local_bundle_adjuster_extended_line::endpoint_trimming (optimize/local_bundle_adjuster_extended_line.cc:676-787) ----
// Restated with the matrix helpers of g2o_lite.hpp (the product's host function in structure-plp-slam_b200/host/
// plpslam_b200_line_trimming.h is written out in scalars).  cam4 = fx fy cx cy.  Returns 1 = keep, 0 = erase.
extern "C" int orc_endpoint_trimming(const double *cam4, const double *pose_cw, const double *plucker, const float *sp,
                                     const float *ep, const double *old_endpoints, double median_depth, double *updated) {
    const Cam c{cam4[0], cam4[1], cam4[2], cam4[3], 0.0};
    Mat3 R;
    for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k) R(r, k) = pose_cw[r * 4 + k];
    const Vec3 t{{pose_cw[3], pose_cw[7], pose_cw[11]}};
    const Vec3 proj = line_project(c, R, t, plucker);
    const double l1 = proj[0], l2 = proj[1], l3 = proj[2];
    double P[12];
    const double K[9] = {c.fx, 0, c.cx, 0, c.fy, c.cy, 0, 0, 1};
    for (int r = 0; r < 3; ++r)
        for (int col = 0; col < 4; ++col) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += K[r * 3 + k] * (col < 3 ? R(k, col) : t[k]);
            P[r * 4 + col] = s;
        }
    const Vec3 m{{plucker[0], plucker[1], plucker[2]}}, d{{plucker[3], plucker[4], plucker[5]}};
    const Mat3 Sm = skew(m);
    const float *pts[2] = {sp, ep};
    for (int e = 0; e < 2; ++e) {
        const double x = pts[e][0], y = pts[e][1];
        const double xc = -(y - (l2 / l1) * x + (l3 / l2)) * ((l1 * l2) / (l1 * l1 + l2 * l2));
        const double yc = -(l1 / l2) * xc - (l3 / l2);
        const double y0 = y - (l2 / l1) * x;
        const Vec3 lt = cross(Vec3{{xc, yc, 1.0}}, Vec3{{0.0, y0, 1.0}});
        double pl[4];
        for (int col = 0; col < 4; ++col) pl[col] = P[col] * lt[0] + P[4 + col] * lt[1] + P[8 + col] * lt[2];
        double X[4];
        for (int r = 0; r < 3; ++r) X[r] = Sm(r, 0) * pl[0] + Sm(r, 1) * pl[1] + Sm(r, 2) * pl[2] + d[r] * pl[3];
        X[3] = -(d[0] * pl[0] + d[1] * pl[1] + d[2] * pl[2]);
        for (int r = 0; r < 3; ++r) updated[3 * e + r] = X[r] / X[3];
    }
    const Vec3 ds{{updated[0] - old_endpoints[0], updated[1] - old_endpoints[1], updated[2] - old_endpoints[2]}};
    const Vec3 de{{updated[3] - old_endpoints[3], updated[4] - old_endpoints[4], updated[5] - old_endpoints[5]}};
    return (norm(ds) / median_depth > 0.1 || norm(de) / median_depth > 0.1) ? 0 : 1;
}
