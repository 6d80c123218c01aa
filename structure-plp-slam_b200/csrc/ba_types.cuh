// ba_types.cuh -- device-side problem descriptor and LM state of the bundle adjuster, shared by the kernels
// (ba_lm_kernels.cuh, ba_chol.cu), the host driver (ba_host.cu) and tests/cta_emu.  No CUDA-runtime dependencies.
#pragma once
#include <stdint.h>

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

}  // namespace plp
