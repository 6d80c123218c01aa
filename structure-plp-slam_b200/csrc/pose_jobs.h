// pose_jobs.h -- device job descriptor of the motion-only BA kernel (plain struct; shared by pose_opt.cu, pipeline.cu and
// the device-code header that tests/cta_emu also compiles for the host).
#pragma once
#include <stdint.h>

#include "../../include/plpslam_b200.h"

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

}  // namespace plp
