// ba_kernels.cuh -- device-side problem descriptor and launch helpers of the local bundle adjuster
// (local_ba.cu = kernels, ba_host.cu = C ABI + LM driver, ba_nccl.cu = multi-GPU collective).
#pragma once
#include "common.cuh"
#include "se3.cuh"

namespace plp {

constexpr int kBaMaxFree = 32;  // non-fixed keyframes whose reduced system (<= 192 x 192) is solved in shared memory; more
                                // keyframes take the dense-in-HBM path (ba_chol.cu)
enum { kBaNeedInit = 0, kBaRunning = 1, kBaDone = 2 };

struct BaState {  // LM state machine, lives in device memory (single writer: the 1-CTA kernels)
    int phase, it, max_it, qmax;
    int iter_start, have_trial, ok2, robust;
    int cur, tries, accepted, solve_active;  // solve_active: the HBM Cholesky kernels of this try have work (large path)
    double lambda, ni, rho, current_chi, scale_pose;
};

struct BaDev {
    // camera
    double fx, fy, cx, cy, bf;
    double delta_pt, delta_ln;  // Huber deltas (sqrt(5.991) | sqrt(7.815), sqrt(5.991))
    // sizes
    int n_kf, n_free, n_pairs, n_pts, n_lines, n_pt_edges, n_ln_edges, n_pl_edges;
    int num_ctas, batch_landmarks, pool_cap, packed_len, packed_sum_len;
    int rank, world;
    int large;            // > kBaMaxFree non-fixed keyframes: reduced system dense in HBM (ba_chol.cu), FP64 atomics
    int phase_init_grid;  // grid of ba_chol_prepare_kernel
    double *dense;        // (6 n_free + 1) x 6 n_free, row-major lower triangle + right-hand side row (large path)
    // keyframes
    const int *kf_hidx;             // index among the free keyframes or -1
    se3::Pose *poses[2];            // current / trial, toggled by BaState::cur
    se3::Pose *pert_pose;           // n_kf x 12: estimate (+)/(-) 1e-9 along each tangent direction
    const int *pair_bi, *pair_bj;   // upper block pairs (bi <= bj) of the reduced camera system
    // landmarks
    double *pts[2];                 // n_pts x 3
    double *lines[2];               // n_lines x 6
    // point edges, grouped by landmark (CSR)
    const int *pt_off, *pt_kf, *pt_lm;
    const float *pt_obs, *pt_info;
    uint8_t *pt_level, *pt_outlier;
    double *pt_chi2, *pt_W;         // last computed chi2; Hpl blocks (24 doubles per edge)
    double *pt_Dinv, *pt_bl;        // per landmark: 16 / 4 doubles
    uint8_t *pt_active;
    const int *pt_plane;            // per point: plane-edge index or -1 (may be null)
    const double *pl_fn;
    double *pl_err;
    // line edges
    const int *ln_off, *ln_kf, *ln_lm;
    const float *ln_obs, *ln_info;
    uint8_t *ln_level, *ln_outlier;
    double *ln_chi2, *ln_W, *ln_Dinv, *ln_bl;
    uint8_t *ln_active;
    // work sharing and reductions
    const int *cta_ranges;          // num_ctas + 1 landmark boundaries (points then lines)
    double *partial;                // num_ctas x packed_len
    double *packed;                 // [S | g | bp | chi | max-diag slots(world)]
    double *dp;                     // 6 x n_free
    double *trial_partial;          // num_ctas x 2
    double *trial_sum;              // 2
    BaState *state;
};

// multi-GPU hook: sum-all-reduce of `n` doubles in place on the context stream (ba_nccl.cu)
struct BaCollective {
    virtual plp_status all_reduce(double *d_buf, int n) = 0;
    virtual bool graph_safe(int n_max) const { return false; }  // may all_reduce(<= n_max doubles) be captured in a CUDA graph?
    virtual void add_calls(uint64_t n) {}                       // a captured graph with n all-reduces was replayed
    virtual plp_status check() { return PLP_OK; }               // after a stream synchronisation
    virtual ~BaCollective() {}
};

size_t ba_linearize_smem(int n_free, int n_pairs, int pool_cap);
int ba_pool_capacity(int n_free, int n_pairs, int max_free_degree);
size_t ba_solve_smem(int n_free);
plp_status ba_prepare_kernels(int n_free, int n_pairs, int pool_cap);
plp_status ba_launch_try(plp_ctx *ctx, const BaDev &B, BaCollective *coll);
size_t ba_dense_bytes(int n_free);
plp_status ba_launch_solve_large(plp_ctx *ctx, const BaDev &B);
plp_status ba_launch_decide(plp_ctx *ctx, const BaDev &B);
plp_status ba_launch_set_state(plp_ctx *ctx, const BaDev &B, int max_it, int robust, int reset_cur);
plp_status ba_launch_classify(plp_ctx *ctx, const BaDev &B, int set_levels);
plp_status ba_launch_init_poses(plp_ctx *ctx, const BaDev &B, const double *d_T_in);
plp_status ba_launch_export(plp_ctx *ctx, const BaDev &B, double *d_T_out, double *d_pts_out, double *d_lines_out);

}  // namespace plp
