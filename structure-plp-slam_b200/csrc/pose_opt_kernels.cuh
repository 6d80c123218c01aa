// pose_opt_kernels.cuh -- device code of the motion-only bundle adjustment (pose_opt.cu launches it).  Free of host-side
// CUDA runtime dependencies so that tests/cta_emu can compile the same text for the host.
//
// optimize::pose_optimizer::optimize (optimize/pose_optimizer.cc:53-229) and
// optimize::pose_optimizer_extended_line::optimize (optimize/pose_optimizer_extended_line.cc:62-305): the g2o graph
// {1 SE3 vertex, one unary edge per matched keypoint / keyline} solved with OptimizationAlgorithmLevenberg, 4 trials x
// <= 10 LM iterations with chi-square re-classification of ALL edges after each trial and removal of the Huber kernels
// at trial 2.
//
// ONE WARP PER FRAME, the whole optimize() in one launch, no block-level barrier anywhere:
//   * lanes stride the edges (FP64 residual + Jacobian, 28 accumulators per lane); the 6 x 6 normal equations are summed
//     with a shuffle butterfly, so EVERY lane ends up with the same H, b and chi2;
//   * every lane then runs the 6 x 6 Cholesky, the SE3 exponential and the LM accept / reject bookkeeping REDUNDANTLY
//     on identical inputs: no broadcast, no barrier, no idle threads waiting for "thread 0" (the first generation -- one
//     256-thread CTA per frame -- spent 43 % of its stall samples in exactly that wait, profiles/source_hotspots_r01j.md);
//   * per-edge state is only the g2o level (= the outlier flag, kept in the caller's output arrays): the chi2 an inlier
//     edge carries into the re-classification is the one of the LAST EVALUATED pose -- also when that step was rejected,
//     g2o's pop() does not recompute (pose_optimizer.cc:177-195) -- and is recomputed from that pose instead of being
//     stored 80 times per call.
// Independent frames are independent warps: a batch of 256 frames is 256 warps in flight at once instead of two waves of
// one-CTA-per-SM blocks.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/plpslam_b200.h"
#include "pose_jobs.h"
#include "se3.cuh"

namespace plp {

namespace po {

constexpr int kWarpsPerCta = 4;
constexpr int kThreads = 32 * kWarpsPerCta;
constexpr int kRed = 28;  // 21 (upper H) + 6 (b) + 1 (chi2)

struct WarpShared {
    se3::Pose pert[12];  // estimate (+)/(-) 1e-9 along each of the 6 tangent directions (numeric line Jacobians)
};

__device__ __forceinline__ double warp_allsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 6x6 SPD solve (Cholesky); H given as upper triangle packed row-wise (21 values); returns false if not SPD
__device__ inline bool solve6(const double *Hu, double lambda, const double *b, double *x) {
    double A[36];
    int k = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) {
            A[i * 6 + j] = Hu[k];
            A[j * 6 + i] = Hu[k];
            ++k;
        }
    for (int i = 0; i < 6; ++i) A[i * 6 + i] += lambda;
    for (int j = 0; j < 6; ++j) {
        double d = A[j * 6 + j];
        for (int q = 0; q < j; ++q) d -= A[j * 6 + q] * A[j * 6 + q];
        if (!(d > 0.0) || !isfinite(d)) return false;
        d = sqrt(d);
        A[j * 6 + j] = d;
        for (int i = j + 1; i < 6; ++i) {
            double s = A[i * 6 + j];
            for (int q = 0; q < j; ++q) s -= A[i * 6 + q] * A[j * 6 + q];
            A[i * 6 + j] = s / d;
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) {
        double s = b[i];
        for (int q = 0; q < i; ++q) s -= A[i * 6 + q] * y[q];
        y[i] = s / A[i * 6 + i];
    }
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
        for (int q = i + 1; q < 6; ++q) s -= A[q * 6 + i] * x[q];
        x[i] = s / A[i * 6 + i];
    }
    return true;
}

__device__ __forceinline__ plp_pt_obs load_pt(const plp_pt_obs *p) {
    plp_pt_obs o;
    o.pos_w[0] = __ldg(&p->pos_w[0]);
    o.pos_w[1] = __ldg(&p->pos_w[1]);
    o.pos_w[2] = __ldg(&p->pos_w[2]);
    o.obs_x = __ldg(&p->obs_x);
    o.obs_y = __ldg(&p->obs_y);
    o.x_right = __ldg(&p->x_right);
    o.inv_sigma_sq = __ldg(&p->inv_sigma_sq);
    return o;
}

__device__ __forceinline__ double eval_point(const se3::Cam &cam, const se3::Pose &P, const plp_pt_obs &o, double *e,
                                             double *pc, bool &stereo) {
    stereo = !(o.x_right < 0);
    se3::map_point(P.R, P.t, o.pos_w, pc);
    const double obs[3] = {(double)o.obs_x, (double)o.obs_y, (double)o.x_right};
    se3::point_error(cam, pc, obs, stereo, e);
    const double w = (double)o.inv_sigma_sq;
    return e[0] * (w * e[0]) + e[1] * (w * e[1]) + (stereo ? e[2] * (w * e[2]) : 0.0);
}

__device__ __forceinline__ double eval_line(const se3::Cam &cam, const se3::Pose &P, const plp_line_obs &o, double *e) {
    const double obs[4] = {(double)o.sp_x, (double)o.sp_y, (double)o.ep_x, (double)o.ep_y};
    se3::line_error(cam, P.R, P.t, o.plucker, obs, e);
    const double w = (double)o.inv_sigma_sq;
    return e[0] * (w * e[0]) + e[1] * (w * e[1]);
}

__global__ void __launch_bounds__(kThreads)
    pose_opt_kernel(const PoseJob *__restrict__ jobs, int batch, plp_camera pcam, plp_pose_opt_cfg cfg) {
    __shared__ WarpShared sh[kWarpsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * kWarpsPerCta + warp;
    if (f >= batch) return;  // a whole warp leaves; nothing below synchronises beyond the warp
    WarpShared &S = sh[warp];
    const PoseJob J = jobs[f];
    const int n_pts = J.n_pts, n_lines = J.n_lines;
    const plp_pt_obs *pts = J.pts;
    const plp_line_obs *lines = J.lines;
    uint8_t *pt_level = J.pt_outlier;    // g2o level of the point edges (1 = outlier) == the output flag
    uint8_t *ln_level = J.line_outlier;  // lane l only ever touches the edges l, l + 32, ...: no cross-lane hazard
    const se3::Cam cam{pcam.fx, pcam.fy, pcam.cx, pcam.cy, pcam.focal_x_baseline};
    // pose_optimizer.cc:120-123: chi-square thresholds (float literals promoted to double)
    const double chi_sq_2D = (double)5.99146f, chi_sq_3D = (double)7.81473f;
    const double delta_pt = pcam.setup_type == 0 ? (double)sqrtf(5.99146f) : (double)sqrtf(7.81473f);
    const double delta_line = (double)sqrtf(5.99146f);

    for (int i = lane; i < n_pts; i += 32) pt_level[i] = 0;
    for (int i = lane; i < n_lines; i += 32) ln_level[i] = 0;
    if (n_pts < 5) {  // pose_optimizer.cc:153-156: nothing is touched
        if (lane < 16) J.T_out[lane] = J.T_in[lane];
        if (lane == 0) {
            *J.n_inliers = 0;
            if (J.lm_iters) *J.lm_iters = 0;
        }
        return;
    }
    se3::Pose est = se3::from_matrix(J.T_in);  // replicated in every lane
    se3::Pose last_eval = est;                 // the pose the inlier edges' errors were last computed at
    bool robust = true;
    int num_bad = 0, lm_iters = 0;
    double lambda = 0, ni = 2;
    for (int trial = 0; trial < cfg.num_trials; ++trial) {
        // ---------------- optimizer.initializeOptimization(); optimizer.optimize(num_each_iter)
        for (int it = 0; it < cfg.num_each_iter; ++it) {
            // perturbed poses for the numeric line Jacobians (BaseUnaryEdge::linearizeOplus, delta = 1e-9)
            if (n_lines > 0) {
                __syncwarp();  // the previous iteration's readers are done
                if (lane < 12) {
                    double u[6] = {0, 0, 0, 0, 0, 0};
                    u[lane >> 1] = (lane & 1) ? -1e-9 : 1e-9;
                    S.pert[lane] = se3::oplus(est, u);
                }
                __syncwarp();
            }
            // computeActiveErrors + buildSystem at the current estimate
            double acc[kRed];
#pragma unroll
            for (int k = 0; k < kRed; ++k) acc[k] = 0;
            for (int i = lane; i < n_pts; i += 32) {
                if (pt_level[i]) continue;
                const plp_pt_obs o = load_pt(pts + i);
                double Jm[18], e[3], pc[3];
                bool stereo;
                const double chi2 = eval_point(cam, est, o, e, pc, stereo);
                se3::point_jac_pose(cam, pc, stereo, Jm);
                double rho0 = chi2, rho1 = 1.0;
                if (robust) se3::huber(chi2, delta_pt, rho0, rho1);
                acc[27] += rho0;
                const double ww = (double)o.inv_sigma_sq * rho1;
                int k = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
#pragma unroll
                    for (int c = a; c < 6; ++c) {
                        acc[k] += ww * (Jm[a] * Jm[c] + Jm[6 + a] * Jm[6 + c] + Jm[12 + a] * Jm[12 + c]);
                        ++k;
                    }
                    acc[21 + a] -= ww * (Jm[a] * e[0] + Jm[6 + a] * e[1] + Jm[12 + a] * e[2]);
                }
            }
            for (int i = lane; i < n_lines; i += 32) {
                if (ln_level[i]) continue;
                const plp_line_obs o = lines[i];
                double Jm[12], e[2];
                const double chi2 = eval_line(cam, est, o, e);
                const double scalar = 1.0 / (2 * 1e-9);
#pragma unroll
                for (int d = 0; d < 6; ++d) {
                    double ep[2], em[2];
                    eval_line(cam, S.pert[2 * d], o, ep);
                    eval_line(cam, S.pert[2 * d + 1], o, em);
                    Jm[d] = scalar * (ep[0] - em[0]);
                    Jm[6 + d] = scalar * (ep[1] - em[1]);
                }
                double rho0 = chi2, rho1 = 1.0;
                if (robust) se3::huber(chi2, delta_line, rho0, rho1);
                acc[27] += rho0;
                const double ww = (double)o.inv_sigma_sq * rho1;
                int k = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
#pragma unroll
                    for (int c = a; c < 6; ++c) {
                        acc[k] += ww * (Jm[a] * Jm[c] + Jm[6 + a] * Jm[6 + c]);
                        ++k;
                    }
                    acc[21 + a] -= ww * (Jm[a] * e[0] + Jm[6 + a] * e[1]);
                }
            }
#pragma unroll
            for (int k = 0; k < kRed; ++k) acc[k] = warp_allsum(acc[k]);
            double current_chi = acc[27];
            if (it == 0) {  // computeLambdaInit: tau * max diag(H)
                double md = 0;
                const int diag[6] = {0, 6, 11, 15, 18, 20};
#pragma unroll
                for (int j = 0; j < 6; ++j) md = fmax(fabs(acc[diag[j]]), md);
                lambda = 1e-5 * md;
                ni = 2;
            }
            int qmax = 0;
            bool terminate = false;
            const double *Hu = acc, *b = acc + 21;
            // ---------------- Levenberg inner loop (<= 10 trials after failure)
            while (true) {
                double x[6] = {0, 0, 0, 0, 0, 0};
                const bool ok2 = solve6(Hu, lambda, b, x);
                const se3::Pose trial_pose = se3::oplus(est, x);
                double chi = 0;
                for (int i = lane; i < n_pts; i += 32) {
                    if (pt_level[i]) continue;
                    const plp_pt_obs o = load_pt(pts + i);
                    double e[3], pc[3];
                    bool stereo;
                    const double chi2 = eval_point(cam, trial_pose, o, e, pc, stereo);
                    double rho0 = chi2, rho1;
                    if (robust) se3::huber(chi2, delta_pt, rho0, rho1);
                    chi += rho0;
                }
                for (int i = lane; i < n_lines; i += 32) {
                    if (ln_level[i]) continue;
                    double e[2];
                    const double chi2 = eval_line(cam, trial_pose, lines[i], e);
                    double rho0 = chi2, rho1;
                    if (robust) se3::huber(chi2, delta_line, rho0, rho1);
                    chi += rho0;
                }
                last_eval = trial_pose;  // the errors stay those of this state even if the step is rejected (g2o pop())
                double temp_chi = warp_allsum(chi);
                if (!ok2) temp_chi = 1.7976931348623157e308;
                double rho = current_chi - temp_chi;
                double scale = 0;
#pragma unroll
                for (int j = 0; j < 6; ++j) scale += x[j] * (lambda * x[j] + b[j]);
                scale += 1e-3;
                rho /= scale;
                bool lambda_finite = true;
                if (rho > 0 && isfinite(temp_chi)) {
                    double alpha = 1. - pow((2 * rho - 1), 3);
                    alpha = fmin(alpha, 2. / 3.);
                    const double sf = fmax(1. / 3., alpha);
                    lambda *= sf;
                    ni = 2;
                    current_chi = temp_chi;
                    est = trial_pose;
                } else {
                    lambda *= ni;
                    ni *= 2;
                    if (!isfinite(lambda)) lambda_finite = false;
                }
                if (lambda_finite) qmax++;
                terminate = (qmax == 10 || rho == 0 || !lambda_finite);
                if (!(lambda_finite && rho < 0 && qmax < 10)) break;
            }
            ++lm_iters;
            if (terminate) break;
        }
        // ---------------- re-classification (pose_optimizer.cc:171-216): inlier edges keep the error of the last
        // evaluated state, outlier edges are recomputed at the current estimate
        int bad = 0;
        for (int i = lane; i < n_pts; i += 32) {
            const plp_pt_obs o = load_pt(pts + i);
            double e[3], pc[3];
            bool stereo;
            const double chi2 = eval_point(cam, pt_level[i] ? est : last_eval, o, e, pc, stereo);
            const bool out = (stereo ? chi_sq_3D : chi_sq_2D) < chi2;
            pt_level[i] = out ? 1 : 0;
            bad += out;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
        num_bad = bad;
        const bool drop_kernel = (trial == cfg.num_trials - 2);
        if (n_pts - num_bad < 5) break;
        for (int i = lane; i < n_lines; i += 32) {  // pose_optimizer_extended_line.cc:269-297
            double e[2];
            const double chi2 = eval_line(cam, ln_level[i] ? est : last_eval, lines[i], e);
            ln_level[i] = chi_sq_2D < chi2 ? 1 : 0;
        }
        if (drop_kernel) robust = false;
    }
    if (lane == 0) {
        double T[16];
        se3::to_matrix(est, T);
        for (int k = 0; k < 16; ++k) J.T_out[k] = T[k];
        *J.n_inliers = n_pts - num_bad;
        if (J.lm_iters) *J.lm_iters = lm_iters;
    }
}

}  // namespace po

}  // namespace plp
