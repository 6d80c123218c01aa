// pose_opt_kernels.cuh -- device code of the motion-only bundle adjustment (pose_opt.cu launches it).  Free of host-side
// CUDA runtime dependencies so that tests/cta_emu can compile the same text for the host.
//
// optimize::pose_optimizer::optimize (optimize/pose_optimizer.cc:53-229) and
// optimize::pose_optimizer_extended_line::optimize (optimize/pose_optimizer_extended_line.cc:62-305): the g2o graph
// {1 SE3 vertex, one unary edge per matched keypoint / keyline} solved with OptimizationAlgorithmLevenberg, 4 trials x
// <= 10 LM iterations with chi-square re-classification of ALL edges after each trial and removal of the Huber kernels
// at trial 2.
//
// ONE 128-THREAD CTA PER FRAME, the whole optimize() in one launch, TWO block barriers per LM try:
//   * threads stride the edges (FP64 residual + Jacobian with ONE reciprocal per edge, 28 accumulators per thread); the
//     6 x 6 normal equations are summed by a shuffle butterfly per warp and a double-buffered shared-memory exchange
//     between the 4 warps (one barrier), so EVERY thread ends up with the same H, b and chi2;
//   * every thread then runs the 6 x 6 Cholesky (fully unrolled, in registers), the SE3 exponential and the LM
//     accept / reject bookkeeping REDUNDANTLY on identical inputs: no broadcast, no "thread 0" section the other threads
//     wait for (the first generation spent 43 % of its stall samples in exactly that wait and nine barriers per try,
//     profiles/source_hotspots_r01j.md; a one-warp-per-frame variant had no barriers at all but only 256 warps in flight
//     for 256 frames: 0.13 instructions per cycle per warp on dependent FP64 divisions, profiles/summary_r02a.md);
//   * per-edge state is only the g2o level (= the outlier flag, kept in the caller's output arrays): the chi2 an inlier
//     edge carries into the re-classification is the one of the LAST EVALUATED pose -- also when that step was rejected,
//     g2o's pop() does not recompute (pose_optimizer.cc:177-195) -- and is recomputed from that pose instead of being
//     stored 80 times per call.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/plpslam_b200.h"
#include "pose_jobs.h"
#include "se3.cuh"

namespace plp {

namespace po {

constexpr int kWarps = 4;              // warps per frame (one CTA = one frame)
constexpr int kThreads = 32 * kWarps;
constexpr int kRed = 28;               // 21 (upper H) + 6 (b) + 1 (chi2)

struct Shared {
    double redH[2][kWarps][kRed];  // per-warp partial sums of (H, b, chi2), double-buffered by LM iteration: ONE barrier per
                                   // reduction, and the totals are re-summed from here where needed instead of living in
                                   // 56 registers across the trial evaluations
    double redc[2][kWarps];        // per-warp partial chi2 of a trial evaluation, double-buffered by try
    se3::Pose pert[12];           // estimate (+)/(-) 1e-9 along each of the 6 tangent directions (numeric line Jacobians)
    se3::Pose last_eval;          // the pose the inlier edges' errors were last computed at (written by thread 0)
    int bad[2][kWarps];
};

__device__ __forceinline__ double warp_allsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sums, EVERY thread receives the totals (fixed order: lanes by butterfly, warps 0..kWarps-1).  One __syncthreads
// per reduction: the buffers alternate, so the readers of reduction n cannot be overtaken by the writers of n + 2 (a thread
// that enters n + 2 has passed the barrier of n + 1, which every reader of n reached after reading).
__device__ __forceinline__ double sumH(const Shared &S, int hp, int k) {
    double a = S.redH[hp][0][k];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) a += S.redH[hp][w][k];
    return a;
}

// 6x6 SPD solve (Cholesky) entirely in registers: fully unrolled, one reciprocal per pivot.  H given as upper triangle
// packed row-wise (21 values) + b, summed from the per-warp partials; returns false if not SPD.
__device__ __forceinline__ bool solve6(const Shared &S, int hp, double lambda, double *x) {
    double A[6][6], b[6];
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) {
                A[j][i] = sumH(S, hp, k);  // lower triangle
                ++k;
            }
#pragma unroll
        for (int i = 0; i < 6; ++i) b[i] = sumH(S, hp, 21 + i);
    }
    double inv[6];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = A[j][j] + lambda;
#pragma unroll
        for (int q = 0; q < j; ++q) d -= A[j][q] * A[j][q];
        if (!(d > 0.0) || !isfinite(d)) ok = false;
        const double r = 1.0 / sqrt(d);  // l_jj = d * r
        inv[j] = r;
        A[j][j] = d * r;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double s = A[i][j];
#pragma unroll
            for (int q = 0; q < j; ++q) s -= A[i][q] * A[j][q];
            A[i][j] = s * r;
        }
    }
    if (!ok) return false;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = b[i];
#pragma unroll
        for (int q = 0; q < i; ++q) s -= A[i][q] * y[q];
        y[i] = s * inv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
#pragma unroll
        for (int q = i + 1; q < 6; ++q) s -= A[q][i] * x[q];
        x[i] = s * inv[i];
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

// Point edge with ONE reciprocal: e = obs - projection (perspective_pose_opt_edge.h:55-113); pc, iz = 1 / z returned for the
// Jacobian.  (The oracle divides by z term by term; the results agree to an ulp or two per term, ~1e-15 on the pose.)
__device__ __forceinline__ double eval_point(const se3::Cam &cam, const se3::Pose &P, const plp_pt_obs &o, double *e,
                                             double *pc, double &iz, bool &stereo) {
    stereo = !(o.x_right < 0);
    se3::map_point(P.R, P.t, o.pos_w, pc);
    iz = 1.0 / pc[2];
    const double rx = cam.fx * pc[0] * iz + cam.cx;
    e[0] = (double)o.obs_x - rx;
    e[1] = (double)o.obs_y - (cam.fy * pc[1] * iz + cam.cy);
    e[2] = stereo ? (double)o.x_right - (rx - cam.bf * iz) : 0.0;
    const double w = (double)o.inv_sigma_sq;
    return e[0] * (w * e[0]) + e[1] * (w * e[1]) + (stereo ? e[2] * (w * e[2]) : 0.0);
}

// d e / d pose (perspective_pose_opt_edge.cc:76-101, :142-173) from pc and iz; J[4] = J[9] = J[16] = 0 are not stored
__device__ __forceinline__ void point_jac(const se3::Cam &c, const double *pc, double iz, bool stereo, double *J /*18*/) {
    const double x = pc[0], y = pc[1], iz2 = iz * iz;
    const double xz = x * iz, yz = y * iz;
    J[0] = xz * yz * c.fx;
    J[1] = -(1.0 + xz * xz) * c.fx;
    J[2] = yz * c.fx;
    J[3] = -iz * c.fx;
    J[4] = 0.0;
    J[5] = x * iz2 * c.fx;
    J[6] = (1.0 + yz * yz) * c.fy;
    J[7] = -xz * yz * c.fy;
    J[8] = -xz * c.fy;
    J[9] = 0.0;
    J[10] = -iz * c.fy;
    J[11] = y * iz2 * c.fy;
    if (stereo) {
        J[12] = J[0] - c.bf * y * iz2;
        J[13] = J[1] + c.bf * x * iz2;
        J[14] = J[2];
        J[15] = J[3];
        J[16] = 0.0;
        J[17] = J[5] - c.bf * iz2;
    }
}

__device__ __forceinline__ double eval_line(const se3::Cam &cam, const se3::Pose &P, const plp_line_obs &o, double *e) {
    const double obs[4] = {(double)o.sp_x, (double)o.sp_y, (double)o.ep_x, (double)o.ep_y};
    se3::line_error(cam, P.R, P.t, o.plucker, obs, e);
    const double w = (double)o.inv_sigma_sq;
    return e[0] * (w * e[0]) + e[1] * (w * e[1]);
}

__global__ void __launch_bounds__(kThreads, 2)
    pose_opt_kernel(const PoseJob *__restrict__ jobs, int batch, plp_camera pcam, plp_pose_opt_cfg cfg) {
    __shared__ Shared S;
    const int tid = threadIdx.x;
    const int f = blockIdx.x;
    if (f >= batch) return;
    const PoseJob J = jobs[f];
    const int n_pts = J.n_pts, n_lines = J.n_lines;
    const plp_pt_obs *pts = J.pts;
    const plp_line_obs *lines = J.lines;
    uint8_t *pt_level = J.pt_outlier;    // g2o level of the point edges (1 = outlier) == the output flag
    uint8_t *ln_level = J.line_outlier;  // thread t only ever touches the edges t, t + kThreads, ...: no cross-thread hazard
    const se3::Cam cam{pcam.fx, pcam.fy, pcam.cx, pcam.cy, pcam.focal_x_baseline};
    // pose_optimizer.cc:120-123: chi-square thresholds (float literals promoted to double)
    const double chi_sq_2D = (double)5.99146f, chi_sq_3D = (double)7.81473f;
    const double delta_pt = pcam.setup_type == 0 ? (double)sqrtf(5.99146f) : (double)sqrtf(7.81473f);
    const double delta_line = (double)sqrtf(5.99146f);
    int hp = 0, cp = 0;  // buffer phases of the H and chi2 reductions

    for (int i = tid; i < n_pts; i += kThreads) pt_level[i] = 0;
    for (int i = tid; i < n_lines; i += kThreads) ln_level[i] = 0;
    if (n_pts < 5) {  // pose_optimizer.cc:153-156: nothing is touched
        if (tid < 16) J.T_out[tid] = J.T_in[tid];
        if (tid == 0) {
            *J.n_inliers = 0;
            if (J.lm_iters) *J.lm_iters = 0;
        }
        return;
    }
    se3::Pose est = se3::from_matrix(J.T_in);  // replicated in every thread
    bool robust = true;
    int num_bad = 0, lm_iters = 0;
    double lambda = 0, ni = 2;
    for (int trial = 0; trial < cfg.num_trials; ++trial) {
        // ---------------- optimizer.initializeOptimization(); optimizer.optimize(num_each_iter)
        for (int it = 0; it < cfg.num_each_iter; ++it) {
            // perturbed poses for the numeric line Jacobians (BaseUnaryEdge::linearizeOplus, delta = 1e-9)
            if (n_lines > 0) {
                __syncthreads();  // the previous iteration's readers are done
                if (tid < 12) {
                    double u[6] = {0, 0, 0, 0, 0, 0};
                    u[tid >> 1] = (tid & 1) ? -1e-9 : 1e-9;
                    S.pert[tid] = se3::oplus(est, u);
                }
                __syncthreads();
            }
            // computeActiveErrors + buildSystem at the current estimate
            double acc[kRed];
#pragma unroll
            for (int k = 0; k < kRed; ++k) acc[k] = 0;
            #pragma unroll 1
            for (int i = tid; i < n_pts; i += kThreads) {
                if (pt_level[i]) continue;
                const plp_pt_obs o = load_pt(pts + i);
                double Jm[18], e[3], pc[3], iz;
                bool stereo;
                const double chi2 = eval_point(cam, est, o, e, pc, iz, stereo);
                point_jac(cam, pc, iz, stereo, Jm);
                double rho0 = chi2, rho1 = 1.0;
                if (robust) se3::huber(chi2, delta_pt, rho0, rho1);
                acc[27] += rho0;
                const double ww = (double)o.inv_sigma_sq * rho1;
                if (!stereo) {  // two rows; J[4] = J[9] = 0 fold away
                    int k = 0;
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        const double wa = ww * Jm[a], wb = ww * Jm[6 + a];
#pragma unroll
                        for (int c = a; c < 6; ++c) {
                            acc[k] += wa * Jm[c] + wb * Jm[6 + c];
                            ++k;
                        }
                        acc[21 + a] -= wa * e[0] + wb * e[1];
                    }
                } else {
                    int k = 0;
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        const double wa = ww * Jm[a], wb = ww * Jm[6 + a], wc = ww * Jm[12 + a];
#pragma unroll
                        for (int c = a; c < 6; ++c) {
                            acc[k] += wa * Jm[c] + wb * Jm[6 + c] + wc * Jm[12 + c];
                            ++k;
                        }
                        acc[21 + a] -= wa * e[0] + wb * e[1] + wc * e[2];
                    }
                }
            }
            #pragma unroll 1
            for (int i = tid; i < n_lines; i += kThreads) {
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
                    const double wa = ww * Jm[a], wb = ww * Jm[6 + a];
#pragma unroll
                    for (int c = a; c < 6; ++c) {
                        acc[k] += wa * Jm[c] + wb * Jm[6 + c];
                        ++k;
                    }
                    acc[21 + a] -= wa * e[0] + wb * e[1];
                }
            }
            hp ^= 1;
#pragma unroll
            for (int k = 0; k < kRed; ++k) {
                const double v = warp_allsum(acc[k]);
                if ((tid & 31) == 0) S.redH[hp][tid >> 5][k] = v;
            }
            __syncthreads();
            double current_chi = sumH(S, hp, 27);
            if (it == 0) {  // computeLambdaInit: tau * max diag(H)
                double md = 0;
                const int diag[6] = {0, 6, 11, 15, 18, 20};
#pragma unroll
                for (int j = 0; j < 6; ++j) md = fmax(fabs(sumH(S, hp, diag[j])), md);
                lambda = 1e-5 * md;
                ni = 2;
            }
            int qmax = 0;
            bool terminate = false;
            // ---------------- Levenberg inner loop (<= 10 trials after failure)
            while (true) {
                double x[6] = {0, 0, 0, 0, 0, 0};
                const bool ok2 = solve6(S, hp, lambda, x);
                const se3::Pose trial_pose = se3::oplus(est, x);
                double chi = 0;
                #pragma unroll 1
                for (int i = tid; i < n_pts; i += kThreads) {
                    if (pt_level[i]) continue;
                    const plp_pt_obs o = load_pt(pts + i);
                    double e[3], pc[3], iz;
                    bool stereo;
                    const double chi2 = eval_point(cam, trial_pose, o, e, pc, iz, stereo);
                    double rho0 = chi2, rho1;
                    if (robust) se3::huber(chi2, delta_pt, rho0, rho1);
                    chi += rho0;
                }
                #pragma unroll 1
                for (int i = tid; i < n_lines; i += kThreads) {
                    if (ln_level[i]) continue;
                    double e[2];
                    const double chi2 = eval_line(cam, trial_pose, lines[i], e);
                    double rho0 = chi2, rho1;
                    if (robust) se3::huber(chi2, delta_line, rho0, rho1);
                    chi += rho0;
                }
                cp ^= 1;
                chi = warp_allsum(chi);
                if ((tid & 31) == 0) S.redc[cp][tid >> 5] = chi;
                __syncthreads();
                // the errors stay those of this state even if the step is rejected (g2o pop()); read after the barrier
                // in front of the re-classification, rewritten only behind the barriers of the next trial's first try
                if (tid == 0) S.last_eval = trial_pose;
                double temp_chi = S.redc[cp][0];
#pragma unroll
                for (int w = 1; w < kWarps; ++w) temp_chi += S.redc[cp][w];
                if (!ok2) temp_chi = 1.7976931348623157e308;
                double rho = current_chi - temp_chi;
                double scale = 0;
#pragma unroll
                for (int j = 0; j < 6; ++j) scale += x[j] * (lambda * x[j] + sumH(S, hp, 21 + j));
                scale += 1e-3;
                rho /= scale;
                bool lambda_finite = true;
                if (rho > 0 && isfinite(temp_chi)) {
                    const double t = 2 * rho - 1;
                    double alpha = 1. - t * t * t;
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
        __syncthreads();  // S.last_eval of the last try is visible
        int bad = 0;
        #pragma unroll 1
        for (int i = tid; i < n_pts; i += kThreads) {
            const plp_pt_obs o = load_pt(pts + i);
            double e[3], pc[3], iz;
            bool stereo;
            const double chi2 = pt_level[i] ? eval_point(cam, est, o, e, pc, iz, stereo)
                                            : eval_point(cam, S.last_eval, o, e, pc, iz, stereo);
            const bool out = (stereo ? chi_sq_3D : chi_sq_2D) < chi2;
            pt_level[i] = out ? 1 : 0;
            bad += out;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
        if ((tid & 31) == 0) S.bad[trial & 1][tid >> 5] = bad;
        __syncthreads();
        num_bad = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) num_bad += S.bad[trial & 1][w];
        const bool drop_kernel = (trial == cfg.num_trials - 2);
        if (n_pts - num_bad < 5) break;
        #pragma unroll 1
        for (int i = tid; i < n_lines; i += kThreads) {  // pose_optimizer_extended_line.cc:269-297
            double e[2];
            const double chi2 = ln_level[i] ? eval_line(cam, est, lines[i], e) : eval_line(cam, S.last_eval, lines[i], e);
            ln_level[i] = chi_sq_2D < chi2 ? 1 : 0;
        }
        if (drop_kernel) robust = false;
    }
    if (tid == 0) {
        double T[16];
        se3::to_matrix(est, T);
        for (int k = 0; k < 16; ++k) J.T_out[k] = T[k];
        *J.n_inliers = n_pts - num_bad;
        if (J.lm_iters) *J.lm_iters = lm_iters;
    }
}

}  // namespace po

}  // namespace plp
