// pose_kernels.cuh -- device job descriptor + launcher of the motion-only BA kernel, shared by pose_opt.cu
// (C ABI) and pipeline.cu (device-resident tracking front-end).
#pragma once
#include "common.cuh"

namespace plp {

struct PoseJob {
    const double *T_in;   // 16
    const plp_pt_obs *pts;
    int n_pts;
    const plp_line_obs *lines;
    int n_lines;
    double *T_out;        // 16
    uint8_t *pt_outlier;
    uint8_t *line_outlier;
    int32_t *n_inliers;
    int32_t *lm_iters;    // may be null
};


plp_status launch_pose_opt(plp_ctx *ctx, const PoseJob *d_jobs, int batch, int max_edges, const plp_camera &cam,
                           const plp_pose_opt_cfg &cfg);

}  // namespace plp
