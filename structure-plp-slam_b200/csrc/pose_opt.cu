// pose_opt.cu -- motion-only bundle adjustment on sm_100a.
//
// Replaces optimize::pose_optimizer::optimize (optimize/pose_optimizer.cc:53-229) and
// optimize::pose_optimizer_extended_line::optimize (optimize/pose_optimizer_extended_line.cc:62-305), i.e. the
// g2o graph {1 SE3 vertex, one unary edge per matched keypoint / keyline} solved with
// OptimizationAlgorithmLevenberg: 4 trials x <= 10 LM iterations with chi-square re-classification of ALL
// edges after each trial and removal of the Huber kernels at trial 2.
//
// One persistent CTA per frame runs the whole optimisation in a single launch (40 LM iterations x launch
// latency would otherwise dominate): threads stride the edges (FP64 residual + Jacobian), the 6x6 normal
// equations are reduced with warp shuffles in a fixed order, thread 0 does the 6x6 Cholesky, the SE3 exp update
// and the LM accept/reject bookkeeping.  The batch dimension is the grid.
//
// g2o semantics that are reproduced on purpose (see DESIGN.md "pose optimiser"):
//   - lambda_0 = 1e-5 * max diag(H) at the first iteration of every optimize() call, nu = 2
//   - rho = (chi_old - chi_new) / (dx^T (lambda dx + b) + 1e-3); accept iff rho > 0 and finite
//   - <= 10 retries per iteration; "terminate" when they are exhausted, rho == 0 or lambda overflows
//   - after a rejected last step the edge errors stay those of the rejected state (g2o does not recompute
//     them on pop()), and the reference thresholds these stale chi2 values for inlier edges
//   - line edges have no analytic Jacobian in the reference -> g2o central differences with delta = 1e-9
#include "common.cuh"
#include "pack.cuh"
#include "se3.cuh"
#include "pose_kernels.cuh"

namespace plp {

namespace {

constexpr int kPoThreads = 256;
constexpr int kPoWarps = kPoThreads / 32;
constexpr int kPoMaxEdges = 6144;  // points + lines per frame (shared-memory bound)
constexpr int kRed = 28;           // 21 (upper H) + 6 (b) + 1 (chi2)

struct PoShared {
    se3::Pose est, trial;
    se3::Pose pert[12];      // est (+)/(-) delta along each of the 6 tangent directions
    double red[kPoWarps][kRed];
    double sum[kRed];
    double x[6];
    double lambda, ni, current_chi, rho;
    int flag_continue, flag_terminate, ok2, qmax;
    int count;
};

__device__ __forceinline__ void block_reduce(PoShared &S, double *v /*kRed per thread*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kRed; ++k) {
        double a = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
        if (lane == 0) S.red[warp][k] = a;
    }
    __syncthreads();
    if (threadIdx.x < kRed) {
        double a = 0;
#pragma unroll
        for (int w = 0; w < kPoWarps; ++w) a += S.red[w][threadIdx.x];
        S.sum[threadIdx.x] = a;
    }
    __syncthreads();
}

// block sum of one value into S.sum[27] with exactly the summation tree of block_reduce (shuffle tree, then warps in order)
__device__ __forceinline__ void block_reduce_chi(PoShared &S, double a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (lane == 0) S.red[warp][27] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < kPoWarps; ++w) t += S.red[w][27];
        S.sum[27] = t;
    }
    __syncthreads();
}

// 6x6 SPD solve (Cholesky); H given as upper triangle packed row-wise (21 values); returns false if not SPD
__device__ bool solve6(const double *Hu, double lambda, const double *b, double *x) {
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

struct EdgeEval {
    double e[3];
    double chi2;
    int dim;
};

__device__ __forceinline__ EdgeEval eval_point(const se3::Cam &cam, const se3::Pose &P, const plp_pt_obs &o, double *pc) {
    EdgeEval r;
    const bool stereo = !(o.x_right < 0);
    se3::map_point(P.R, P.t, o.pos_w, pc);
    const double obs[3] = {(double)o.obs_x, (double)o.obs_y, (double)o.x_right};
    se3::point_error(cam, pc, obs, stereo, r.e);
    const double w = (double)o.inv_sigma_sq;
    r.dim = stereo ? 3 : 2;
    r.chi2 = r.e[0] * (w * r.e[0]) + r.e[1] * (w * r.e[1]) + (stereo ? r.e[2] * (w * r.e[2]) : 0.0);
    return r;
}

__device__ __forceinline__ EdgeEval eval_line(const se3::Cam &cam, const se3::Pose &P, const plp_line_obs &o) {
    EdgeEval r;
    const double obs[4] = {(double)o.sp_x, (double)o.sp_y, (double)o.ep_x, (double)o.ep_y};
    se3::line_error(cam, P.R, P.t, o.plucker, obs, r.e);
    r.e[2] = 0;
    const double w = (double)o.inv_sigma_sq;
    r.dim = 2;
    r.chi2 = r.e[0] * (w * r.e[0]) + r.e[1] * (w * r.e[1]);
    return r;
}

__global__ void __launch_bounds__(kPoThreads, 1)
    pose_opt_kernel(const PoseJob *__restrict__ jobs, plp_camera pcam, plp_pose_opt_cfg cfg, int stage_cap) {
    extern __shared__ __align__(16) uint8_t po_smem[];
    PoShared &S = *reinterpret_cast<PoShared *>(po_smem);
    const PoseJob J = jobs[blockIdx.x];
    const int n_pts = J.n_pts, n_lines = J.n_lines, n_edges = n_pts + n_lines;
    double *chi2_last = reinterpret_cast<double *>(po_smem + ((sizeof(PoShared) + 15) & ~(size_t)15));
    uint8_t *level = reinterpret_cast<uint8_t *>(chi2_last + n_edges);  // 1 = outlier (g2o level 1)
    const int tid = threadIdx.x;
    // The observations are read twice per LM iteration (system build + trial evaluation), ~80 times per call: stage
    // them in shared memory once when they fit (they do at the config sizes), else read them through L2.
    const plp_pt_obs *pts = J.pts;
    const plp_line_obs *lines = J.lines;
    {
        uint8_t *stage = po_smem + ((((sizeof(PoShared) + 15) & ~(size_t)15) + (size_t)n_edges * 9 + 15) & ~(size_t)15);
        const size_t pt_bytes = (size_t)n_pts * sizeof(plp_pt_obs), ln_bytes = (size_t)n_lines * sizeof(plp_line_obs);
        if (pt_bytes + ln_bytes <= (size_t)stage_cap) {
            static_assert(sizeof(plp_pt_obs) % 8 == 0 && sizeof(plp_line_obs) % 8 == 0, "observation PODs are 8-byte multiples");
            double *dst = reinterpret_cast<double *>(stage);
            const double *src_p = reinterpret_cast<const double *>(J.pts), *src_l = reinterpret_cast<const double *>(J.lines);
            const int wp = (int)(pt_bytes / 8), wl = (int)(ln_bytes / 8);
            for (int i = tid; i < wp; i += kPoThreads) dst[i] = src_p[i];
            for (int i = tid; i < wl; i += kPoThreads) dst[wp + i] = src_l[i];
            pts = reinterpret_cast<const plp_pt_obs *>(stage);
            lines = reinterpret_cast<const plp_line_obs *>(stage + pt_bytes);
        }
    }
    const se3::Cam cam{pcam.fx, pcam.fy, pcam.cx, pcam.cy, pcam.focal_x_baseline};
    // pose_optimizer.cc:120-123: chi-square thresholds (float literals promoted to double)
    const double chi_sq_2D = (double)5.99146f, chi_sq_3D = (double)7.81473f;
    const double delta_pt = pcam.setup_type == 0 ? (double)sqrtf(5.99146f) : (double)sqrtf(7.81473f);
    const double delta_line = (double)sqrtf(5.99146f);

    for (int i = tid; i < n_edges; i += kPoThreads) {
        level[i] = 0;
        chi2_last[i] = 0.0;
    }
    for (int i = tid; i < n_pts; i += kPoThreads) J.pt_outlier[i] = 0;
    for (int i = tid; i < n_lines; i += kPoThreads) J.line_outlier[i] = 0;
    if (tid == 0) {
        S.est = se3::from_matrix(J.T_in);
        S.count = 0;
    }
    __syncthreads();
    if (n_pts < 5) {  // pose_optimizer.cc:153-156: nothing is touched
        if (tid < 16) J.T_out[tid] = J.T_in[tid];
        if (tid == 0) {
            *J.n_inliers = 0;
            if (J.lm_iters) *J.lm_iters = 0;
        }
        return;
    }
    bool robust = true;
    int num_bad = 0, lm_iters = 0;
    for (int trial = 0; trial < cfg.num_trials; ++trial) {
        // ---------------- optimizer.initializeOptimization(); optimizer.optimize(num_each_iter)
        for (int it = 0; it < cfg.num_each_iter; ++it) {
            // perturbed poses for the numeric line Jacobians (BaseUnaryEdge::linearizeOplus, delta = 1e-9)
            if (n_lines > 0 && tid < 12) {
                double u[6] = {0, 0, 0, 0, 0, 0};
                u[tid >> 1] = (tid & 1) ? -1e-9 : 1e-9;
                S.pert[tid] = se3::oplus(S.est, u);
            }
            __syncthreads();
            // computeActiveErrors + buildSystem at the current estimate
            double acc[kRed];
#pragma unroll
            for (int k = 0; k < kRed; ++k) acc[k] = 0;
            for (int i = tid; i < n_edges; i += kPoThreads) {
                if (level[i]) continue;
                double Jm[18];
                EdgeEval ev;
                double w, delta;
                if (i < n_pts) {
                    const plp_pt_obs o = pts[i];
                    double pc[3];
                    ev = eval_point(cam, S.est, o, pc);
                    se3::point_jac_pose(cam, pc, ev.dim == 3, Jm);
                    w = (double)o.inv_sigma_sq;
                    delta = delta_pt;
                } else {
                    const plp_line_obs o = lines[i - n_pts];
                    ev = eval_line(cam, S.est, o);
                    const double scalar = 1.0 / (2 * 1e-9);
#pragma unroll
                    for (int d = 0; d < 6; ++d) {
                        const EdgeEval ep = eval_line(cam, S.pert[2 * d], o), em = eval_line(cam, S.pert[2 * d + 1], o);
                        Jm[d] = scalar * (ep.e[0] - em.e[0]);
                        Jm[6 + d] = scalar * (ep.e[1] - em.e[1]);
                        Jm[12 + d] = 0;
                    }
                    w = (double)o.inv_sigma_sq;
                    delta = delta_line;
                }
                chi2_last[i] = ev.chi2;
                double rho0 = ev.chi2, rho1 = 1.0;
                if (robust) se3::huber(ev.chi2, delta, rho0, rho1);
                acc[27] += rho0;
                const double ww = w * rho1;
                int k = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
#pragma unroll
                    for (int c = a; c < 6; ++c) {
                        acc[k] += ww * (Jm[a] * Jm[c] + Jm[6 + a] * Jm[6 + c] + Jm[12 + a] * Jm[12 + c]);
                        ++k;
                    }
                    acc[21 + a] -= ww * (Jm[a] * ev.e[0] + Jm[6 + a] * ev.e[1] + Jm[12 + a] * ev.e[2]);
                }
            }
            block_reduce(S, acc);
            if (tid == 0) {
                S.current_chi = S.sum[27];
                if (it == 0) {  // computeLambdaInit: tau * max diag(H)
                    double md = 0;
                    const int diag[6] = {0, 6, 11, 15, 18, 20};
                    for (int j = 0; j < 6; ++j) md = fmax(fabs(S.sum[diag[j]]), md);
                    S.lambda = 1e-5 * md;
                    S.ni = 2;
                }
                S.qmax = 0;
                S.flag_terminate = 0;
            }
            __syncthreads();
            double Hu[21], b[6];
#pragma unroll
            for (int k = 0; k < 21; ++k) Hu[k] = S.sum[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) b[k] = S.sum[21 + k];
            // ---------------- Levenberg inner loop (<= 10 trials after failure)
            while (true) {
                if (tid == 0) {
                    double x[6] = {0, 0, 0, 0, 0, 0};
                    S.ok2 = solve6(Hu, S.lambda, b, x) ? 1 : 0;
                    for (int k = 0; k < 6; ++k) S.x[k] = x[k];
                    S.trial = se3::oplus(S.est, x);
                }
                __syncthreads();
                double chi = 0;
                for (int i = tid; i < n_edges; i += kPoThreads) {
                    if (level[i]) continue;
                    EdgeEval ev;
                    double delta;
                    if (i < n_pts) {
                        double pc[3];
                        ev = eval_point(cam, S.trial, pts[i], pc);
                        delta = delta_pt;
                    } else {
                        ev = eval_line(cam, S.trial, lines[i - n_pts]);
                        delta = delta_line;
                    }
                    chi2_last[i] = ev.chi2;  // stays even if the step is rejected (g2o pop() does not recompute)
                    double rho0 = ev.chi2, rho1;
                    if (robust) se3::huber(ev.chi2, delta, rho0, rho1);
                    chi += rho0;
                }
                block_reduce_chi(S, chi);
                if (tid == 0) {
                    double temp_chi = S.sum[27];
                    if (!S.ok2) temp_chi = 1.7976931348623157e308;
                    double rho = S.current_chi - temp_chi;
                    double scale = 0;
                    for (int j = 0; j < 6; ++j) scale += S.x[j] * (S.lambda * S.x[j] + b[j]);
                    scale += 1e-3;
                    rho /= scale;
                    bool lambda_finite = true;
                    if (rho > 0 && isfinite(temp_chi)) {
                        double alpha = 1. - pow((2 * rho - 1), 3);
                        alpha = fmin(alpha, 2. / 3.);
                        const double sf = fmax(1. / 3., alpha);
                        S.lambda *= sf;
                        S.ni = 2;
                        S.current_chi = temp_chi;
                        S.est = S.trial;
                    } else {
                        S.lambda *= S.ni;
                        S.ni *= 2;
                        if (!isfinite(S.lambda)) lambda_finite = false;
                    }
                    if (lambda_finite) S.qmax++;
                    S.rho = rho;
                    S.flag_continue = (lambda_finite && rho < 0 && S.qmax < 10) ? 1 : 0;
                    S.flag_terminate = (S.qmax == 10 || rho == 0 || !lambda_finite) ? 1 : 0;
                }
                __syncthreads();
                if (!S.flag_continue) break;
            }
            ++lm_iters;
            if (S.flag_terminate) break;
            __syncthreads();
        }
        __syncthreads();
        // ---------------- re-classification (pose_optimizer.cc:171-216)
        int bad = 0;
        for (int i = tid; i < n_pts; i += kPoThreads) {
            const plp_pt_obs o = pts[i];
            double chi2 = chi2_last[i];
            if (level[i]) {  // outlier edges are recomputed at the current estimate
                double pc[3];
                chi2 = eval_point(cam, S.est, o, pc).chi2;
                chi2_last[i] = chi2;
            }
            const bool mono = o.x_right < 0;
            const bool out = (mono ? chi_sq_2D : chi_sq_3D) < chi2;
            level[i] = out ? 1 : 0;
            J.pt_outlier[i] = out ? 1 : 0;
            bad += out;
        }
        num_bad = __syncthreads_count(0);  // barrier; real count below
        {
            // block-wide sum of `bad`
            int v = bad;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0) atomicAdd(&S.count, v);
            __syncthreads();
            num_bad = S.count;
            __syncthreads();
            if (tid == 0) S.count = 0;
        }
        const bool drop_kernel = (trial == cfg.num_trials - 2);
        if (n_pts - num_bad < 5) {
            if (drop_kernel) robust = false;
            break;
        }
        for (int i = tid; i < n_lines; i += kPoThreads) {  // pose_optimizer_extended_line.cc:269-297
            const int e = n_pts + i;
            double chi2 = chi2_last[e];
            if (level[e]) {
                chi2 = eval_line(cam, S.est, lines[i]).chi2;
                chi2_last[e] = chi2;
            }
            const bool out = chi_sq_2D < chi2;
            level[e] = out ? 1 : 0;
            J.line_outlier[i] = out ? 1 : 0;
        }
        if (drop_kernel) robust = false;
        __syncthreads();
    }
    if (tid == 0) {
        double T[16];
        se3::to_matrix(S.est, T);
        for (int k = 0; k < 16; ++k) J.T_out[k] = T[k];
        *J.n_inliers = n_pts - num_bad;
        if (J.lm_iters) *J.lm_iters = lm_iters;
    }
}

__global__ void build_pose_jobs_kernel(PoseJob *jobs, int batch, const double *T_in, const plp_pt_obs *pts,
                                       const int32_t *pt_off, const plp_line_obs *lines, const int32_t *line_off,
                                       double *T_out, uint8_t *pt_outlier, uint8_t *line_outlier, int32_t *n_inl,
                                       int32_t *lm_iters) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    PoseJob J;
    J.T_in = T_in + 16 * (size_t)b;
    J.pts = pts + pt_off[b];
    J.n_pts = pt_off[b + 1] - pt_off[b];
    const int l0 = line_off ? line_off[b] : 0, l1 = line_off ? line_off[b + 1] : 0;
    J.lines = lines ? lines + l0 : nullptr;
    J.n_lines = lines ? l1 - l0 : 0;
    J.T_out = T_out + 16 * (size_t)b;
    J.pt_outlier = pt_outlier + pt_off[b];
    J.line_outlier = line_outlier ? line_outlier + l0 : nullptr;
    J.n_inliers = n_inl + b;
    J.lm_iters = lm_iters ? lm_iters + b : nullptr;
    jobs[b] = J;
}

}  // namespace

// staging area for the observations: every edge could be a line (72 B), capped so that the CTA stays within 200 KB
static size_t pose_stage_bytes(int max_edges) {
    return std::min((size_t)max_edges * sizeof(plp_line_obs), (size_t)144 * 1024);
}
size_t pose_smem_bytes(int max_edges) {
    return ((sizeof(PoShared) + 15) & ~(size_t)15) + (size_t)max_edges * 9 + 64 + pose_stage_bytes(max_edges);
}

plp_status launch_pose_opt(plp_ctx *ctx, const PoseJob *d_jobs, int batch, int max_edges, const plp_camera &cam,
                           const plp_pose_opt_cfg &cfg) {
    if (batch <= 0) return PLP_OK;
    if (max_edges > kPoMaxEdges) {
        set_error("pose optimiser: %d edges exceed the per-frame capacity %d", max_edges, kPoMaxEdges);
        return PLP_ERR_CAPACITY;
    }
    const size_t smem = pose_smem_bytes(max_edges < 64 ? 64 : max_edges);
    PLP_SMEM_OPTIN(pose_opt_kernel, smem);
    PLP_LAUNCH(ctx, pose_opt_kernel, batch, kPoThreads, smem, d_jobs, cam, cfg,
               (int)pose_stage_bytes(max_edges < 64 ? 64 : max_edges));
    PLP_CHECK_LAUNCH();
    return PLP_OK;
}


}  // namespace plp

using namespace plp;

extern "C" {

plp_status plp_pose_optimize_batch_dev(plp_ctx *ctx, const plp_camera *cam, int batch, const double *d_T_in,
                                       const plp_pt_obs *d_pts, const int32_t *d_pt_off, const plp_line_obs *d_lines,
                                       const int32_t *d_line_off, int max_edges_per_frame, const plp_pose_opt_cfg *cfg,
                                       double *d_T_out, uint8_t *d_pt_outlier, uint8_t *d_line_outlier,
                                       int32_t *d_n_inliers, int32_t *d_lm_iters) {
    PLP_REQUIRE(ctx && cam && cfg && d_T_in && d_pts && d_pt_off && d_T_out && d_pt_outlier && d_n_inliers,
                "null pointer");
    PLP_REQUIRE(batch >= 0 && cfg->num_trials >= 1 && cfg->num_each_iter >= 1, "batch / cfg");
    PLP_REQUIRE(!d_lines || (d_line_off && d_line_outlier), "line arrays");
    if (batch == 0) return PLP_OK;
    PLP_CUDA_TRY(cudaSetDevice(ctx->device));
    void *d_jobs = nullptr;
    PLP_TRY(ctx_scratch(ctx, 1, (size_t)batch * sizeof(PoseJob), &d_jobs));
    PLP_LAUNCH(ctx, build_pose_jobs_kernel, div_up(batch, 128), 128, 0, (PoseJob *)d_jobs, batch, d_T_in, d_pts,
               d_pt_off, d_lines, d_line_off, d_T_out, d_pt_outlier, d_line_outlier, d_n_inliers, d_lm_iters);
    PLP_CHECK_LAUNCH();
    return launch_pose_opt(ctx, (const PoseJob *)d_jobs, batch, max_edges_per_frame, *cam, *cfg);
}

plp_status plp_pose_optimize_batch(plp_ctx *ctx, const plp_camera *cam, int batch, const double *T_in,
                                   const plp_pt_obs *pts, const int32_t *pt_off, const plp_line_obs *lines,
                                   const int32_t *line_off, const plp_pose_opt_cfg *cfg, double *T_out,
                                   uint8_t *pt_outlier, uint8_t *line_outlier, int32_t *n_inliers) {
    PLP_REQUIRE(ctx && cam && cfg && T_in && pt_off && T_out && n_inliers, "null pointer");
    PLP_REQUIRE(batch >= 0, "batch");
    if (batch == 0) return PLP_OK;
    const int n_pts = pt_off[batch], n_lines = line_off ? line_off[batch] : 0;
    PLP_REQUIRE(n_pts >= 0 && n_lines >= 0 && pt_off[0] == 0 && (!line_off || line_off[0] == 0), "offsets");
    PLP_REQUIRE((n_pts == 0 || (pts && pt_outlier)) && (n_lines == 0 || (lines && line_outlier)), "null arrays");
    int max_edges = 0;
    for (int b = 0; b < batch; ++b) {
        const int e = (pt_off[b + 1] - pt_off[b]) + (line_off ? line_off[b + 1] - line_off[b] : 0);
        PLP_REQUIRE(pt_off[b + 1] >= pt_off[b] && (!line_off || line_off[b + 1] >= line_off[b]), "offsets not monotone");
        max_edges = e > max_edges ? e : max_edges;
    }
    PLP_CUDA_TRY(cudaSetDevice(ctx->device));
    Packer pk;
    std::vector<int32_t> zero_off(batch + 1, 0);
    const size_t o_T = pk.add(T_in, (size_t)batch * 128);
    const size_t o_pts = pk.add(n_pts ? (const void *)pts : (const void *)zero_off.data(), n_pts ? (size_t)n_pts * sizeof(plp_pt_obs) : 8);
    const size_t o_po = pk.add(pt_off, (size_t)(batch + 1) * 4);
    const size_t o_lines = n_lines ? pk.add(lines, (size_t)n_lines * sizeof(plp_line_obs)) : Packer::kNone;
    const size_t o_lo = n_lines ? pk.add(line_off, (size_t)(batch + 1) * 4) : Packer::kNone;
    const size_t o_To = pk.reserve((size_t)batch * 128);
    const size_t o_pout = pk.reserve((size_t)n_pts + 8), o_lout = pk.reserve((size_t)n_lines + 8);
    const size_t o_inl = pk.reserve((size_t)batch * 4);
    uint8_t *d;
    PLP_TRY(pk.upload(ctx, 0, &d));
    PLP_TRY(plp_pose_optimize_batch_dev(ctx, cam, batch, Packer::at<double>(d, o_T), Packer::at<plp_pt_obs>(d, o_pts),
                                        Packer::at<int32_t>(d, o_po), Packer::at<plp_line_obs>(d, o_lines),
                                        Packer::at<int32_t>(d, o_lo), max_edges, cfg, Packer::at<double>(d, o_To),
                                        Packer::at<uint8_t>(d, o_pout), n_lines ? Packer::at<uint8_t>(d, o_lout) : nullptr,
                                        Packer::at<int32_t>(d, o_inl), nullptr));
    PLP_CUDA_TRY(cudaMemcpyAsync(T_out, d + o_To, (size_t)batch * 128, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_pts) PLP_CUDA_TRY(cudaMemcpyAsync(pt_outlier, d + o_pout, (size_t)n_pts, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_lines) PLP_CUDA_TRY(cudaMemcpyAsync(line_outlier, d + o_lout, (size_t)n_lines, cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaMemcpyAsync(n_inliers, d + o_inl, (size_t)batch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PLP_OK;
}

plp_status plp_pose_optimize(plp_ctx *ctx, const plp_camera *cam, const double *T_cw_in, const plp_pt_obs *pts,
                             int n_pts, const plp_line_obs *lines, int n_lines, const plp_pose_opt_cfg *cfg,
                             double *T_cw_out, uint8_t *pt_outlier, uint8_t *line_outlier, int32_t *n_inliers_out) {
    PLP_REQUIRE(n_pts >= 0 && n_lines >= 0, "sizes");
    const int32_t po[2] = {0, n_pts}, lo[2] = {0, n_lines};
    return plp_pose_optimize_batch(ctx, cam, 1, T_cw_in, pts, po, n_lines ? lines : nullptr, n_lines ? lo : nullptr, cfg,
                                   T_cw_out, pt_outlier, line_outlier, n_inliers_out);
}

}  // extern "C"
