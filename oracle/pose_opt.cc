// oracle/pose_opt.cc -- CPU restatement of optimize::pose_optimizer and pose_optimizer_extended_line
// (TEST INFRASTRUCTURE ONLY).  Follows optimize/pose_optimizer.cc:53-229 and
// optimize/pose_optimizer_extended_line.cc:62-305; the g2o machinery is restated in g2o_lite.hpp
// (PARITY UNPINNED, see there).
#include <algorithm>
#include <cstdint>

#include "g2o_lite.hpp"
#include "pose_opt.h"

using namespace g2o_lite;

namespace {

struct Edge {
    bool is_line = false;
    bool stereo = false;  // 3-D error (points with x_right >= 0)
    Vec3 Xw{{0, 0, 0}};
    double plucker[6] = {0, 0, 0, 0, 0, 0};
    double obs[4] = {0, 0, 0, 0};
    double info = 1.0;    // information = info * I
    bool robust = true;   // robust kernel still attached
    int level = 0;
    double err[3] = {0, 0, 0};  // _error as left by the last computeError()
    int dim() const { return is_line ? 2 : (stereo ? 3 : 2); }
    double chi2() const {  // e^T Omega e
        double s = 0;
        for (int i = 0; i < dim(); ++i) s += err[i] * (info * err[i]);
        return s;
    }
};

struct Problem {
    Cam cam;
    SE3 est;
    std::vector<Edge> edges;
    double delta_pt = 0, delta_line = 0;

    void compute_error(Edge &e, const SE3 &pose) const {
        const Mat3 R = pose.R();
        if (e.is_line)
            line_error(cam, R, pose.t, e.plucker, e.obs, e.err);
        else
            point_error(cam, R, pose.t, e.Xw, e.obs, e.stereo, e.err);
    }
    // linearizeOplus: analytic for points, g2o's central-difference numeric Jacobian for lines
    void jacobian(Edge &e, double *J /*dim x 6*/) const {
        if (!e.is_line) {
            const Vec3 pc = est.R() * e.Xw + est.t;
            point_jac_pose(cam, pc, e.stereo, J);
            return;
        }
        const double delta = 1e-9, scalar = 1 / (2 * delta);
        double before[3] = {e.err[0], e.err[1], e.err[2]};
        for (int d = 0; d < 6; ++d) {
            double add[6] = {0, 0, 0, 0, 0, 0};
            add[d] = delta;
            Edge tmp = e;
            compute_error(tmp, se3_oplus(est, add));
            const double e1[2] = {tmp.err[0], tmp.err[1]};
            add[d] = -delta;
            compute_error(tmp, se3_oplus(est, add));
            J[0 * 6 + d] = scalar * (e1[0] - tmp.err[0]);
            J[1 * 6 + d] = scalar * (e1[1] - tmp.err[1]);
        }
        e.err[0] = before[0];
        e.err[1] = before[1];
        e.err[2] = before[2];
    }
    double delta_of(const Edge &e) const { return e.is_line ? delta_line : delta_pt; }

    void compute_active_errors(const std::vector<int> &active) {
        for (int i : active) compute_error(edges[i], est);
    }
    double active_robust_chi2(const std::vector<int> &active) const {
        double chi = 0;
        for (int i : active) {
            const Edge &e = edges[i];
            if (e.robust) {
                double rho[3];
                huber(e.chi2(), delta_of(e), rho);
                chi += rho[0];
            } else {
                chi += e.chi2();
            }
        }
        return chi;
    }
    void build_system(const std::vector<int> &active, double *H /*36*/, double *b /*6*/) {
        std::fill(H, H + 36, 0.0);
        std::fill(b, b + 6, 0.0);
        for (int i : active) {
            Edge &e = edges[i];
            double J[18];
            jacobian(e, J);
            const int D = e.dim();
            double w = e.info;
            if (e.robust) {
                double rho[3];
                huber(e.chi2(), delta_of(e), rho);
                w *= rho[1];
            }
            // b += J^T (-w e) ; H += J^T w J
            for (int r = 0; r < D; ++r) {
                const double we = -w * e.err[r];
                for (int a = 0; a < 6; ++a) {
                    b[a] += J[r * 6 + a] * we;
                    for (int c = 0; c < 6; ++c) H[a * 6 + c] += J[r * 6 + a] * w * J[r * 6 + c];
                }
            }
        }
    }

    // SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg
    int optimize(const std::vector<int> &active, int iterations, int *lm_iterations_total) {
        double lambda = 0, ni = 2;
        int done = 0;
        for (int it = 0; it < iterations; ++it) {
            compute_active_errors(active);
            double current_chi = active_robust_chi2(active);
            double temp_chi = current_chi;
            double H[36], b[6];
            build_system(active, H, b);
            if (it == 0) {  // computeLambdaInit
                double max_diag = 0;
                for (int j = 0; j < 6; ++j) max_diag = std::max(std::fabs(H[j * 6 + j]), max_diag);
                lambda = 1e-5 * max_diag;
                ni = 2;
            }
            double rho = 0;
            int qmax = 0;
            bool lambda_finite = true;
            do {
                const SE3 backup = est;  // push
                std::vector<double> Hl(H, H + 36);
                for (int j = 0; j < 6; ++j) Hl[j * 6 + j] += lambda;
                double x[6] = {0, 0, 0, 0, 0, 0};
                const bool ok2 = cholesky_solve(Hl, b, x, 6);
                est = se3_oplus(est, x);
                compute_active_errors(active);
                temp_chi = active_robust_chi2(active);
                if (!ok2) temp_chi = std::numeric_limits<double>::max();
                rho = current_chi - temp_chi;
                double scale = 0;
                for (int j = 0; j < 6; ++j) scale += x[j] * (lambda * x[j] + b[j]);
                scale += 1e-3;
                rho /= scale;
                if (rho > 0 && std::isfinite(temp_chi)) {
                    double alpha = 1. - std::pow((2 * rho - 1), 3);
                    alpha = std::min(alpha, 2. / 3.);
                    const double scale_factor = std::max(1. / 3., alpha);
                    lambda *= scale_factor;
                    ni = 2;
                    current_chi = temp_chi;
                } else {
                    lambda *= ni;
                    ni *= 2;
                    est = backup;  // pop (the edge errors are NOT recomputed, as in g2o)
                    if (!std::isfinite(lambda)) {
                        lambda_finite = false;
                        break;
                    }
                }
                qmax++;
            } while (rho < 0 && qmax < 10);
            ++done;
            if (lm_iterations_total) ++*lm_iterations_total;
            if (qmax == 10 || rho == 0 || !lambda_finite) break;  // Terminate
        }
        return done;
    }
};

}  // namespace

extern "C" int orc_pose_optimize(const orc_pose_cam *cam, const double *T_cw_in, const orc_pt_obs *pts, int n_pts,
                                 const orc_line_obs *lines, int n_lines, int num_trials, int num_each_iter,
                                 double *T_cw_out, uint8_t *pt_outlier, uint8_t *line_outlier,
                                 int *lm_iterations_out) {
    // pose_optimizer.cc:120-156 / pose_optimizer_extended_line.cc:110-188
    const float chi_sq_2D = 5.99146f, chi_sq_3D = 7.81473f;
    const float sqrt_chi_sq_2D = std::sqrt(chi_sq_2D), sqrt_chi_sq_3D = std::sqrt(chi_sq_3D);
    if (lm_iterations_out) *lm_iterations_out = 0;
    for (int i = 0; i < 16; ++i) T_cw_out[i] = T_cw_in[i];
    const unsigned num_init_obs = (unsigned)n_pts;
    for (int i = 0; i < n_pts; ++i) pt_outlier[i] = 0;
    if (num_init_obs < 5) return 0;
    Problem P;
    P.cam = {cam->fx, cam->fy, cam->cx, cam->cy, cam->focal_x_baseline};
    P.est = se3_from_matrix(T_cw_in);
    P.delta_pt = cam->setup_type == 0 ? sqrt_chi_sq_2D : sqrt_chi_sq_3D;
    P.delta_line = sqrt_chi_sq_2D;
    P.edges.resize(n_pts + n_lines);
    for (int i = 0; i < n_pts; ++i) {
        Edge &e = P.edges[i];
        e.stereo = !(pts[i].x_right < 0);  // is_monocular_ = obs_x_right < 0
        e.Xw = {{pts[i].pos_w[0], pts[i].pos_w[1], pts[i].pos_w[2]}};
        e.obs[0] = pts[i].obs_x;
        e.obs[1] = pts[i].obs_y;
        e.obs[2] = pts[i].x_right;
        e.info = pts[i].inv_sigma_sq;
    }
    for (int i = 0; i < n_lines; ++i) {
        Edge &e = P.edges[n_pts + i];
        e.is_line = true;
        for (int k = 0; k < 6; ++k) e.plucker[k] = lines[i].plucker[k];
        e.obs[0] = lines[i].sp_x;
        e.obs[1] = lines[i].sp_y;
        e.obs[2] = lines[i].ep_x;
        e.obs[3] = lines[i].ep_y;
        e.info = lines[i].inv_sigma_sq;
        line_outlier[i] = 0;
    }
    unsigned num_bad_obs = 0;
    for (int trial = 0; trial < num_trials; ++trial) {
        // initializeOptimization(): level-0 edges
        std::vector<int> active;
        for (int i = 0; i < (int)P.edges.size(); ++i)
            if (P.edges[i].level == 0) active.push_back(i);
        P.optimize(active, num_each_iter, lm_iterations_out);
        num_bad_obs = 0;
        for (int i = 0; i < n_pts; ++i) {  // pose_optimizer.cc:171-216
            Edge &e = P.edges[i];
            if (pt_outlier[i]) P.compute_error(e, P.est);
            const double thr = e.stereo ? (double)chi_sq_3D : (double)chi_sq_2D;
            if (thr < e.chi2()) {
                pt_outlier[i] = 1;
                e.level = 1;
                ++num_bad_obs;
            } else {
                pt_outlier[i] = 0;
                e.level = 0;
            }
            if (trial == num_trials - 2) e.robust = false;
        }
        if (num_init_obs - num_bad_obs < 5) break;
        for (int i = 0; i < n_lines; ++i) {  // pose_optimizer_extended_line.cc:269-297
            Edge &e = P.edges[n_pts + i];
            if (line_outlier[i]) P.compute_error(e, P.est);
            if ((double)chi_sq_2D < e.chi2()) {
                line_outlier[i] = 1;
                e.level = 1;
            } else {
                line_outlier[i] = 0;
                e.level = 0;
            }
            if (trial == num_trials - 2) e.robust = false;
        }
    }
    se3_to_matrix(P.est, T_cw_out);
    return (int)(num_init_obs - num_bad_obs);
}
