/* oracle/pose_opt.h -- pose optimiser / local BA restatement (TEST INFRASTRUCTURE ONLY). */
#ifndef PLP_ORACLE_POSE_OPT_H
#define PLP_ORACLE_POSE_OPT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_pose_cam {
    double fx, fy, cx, cy;
    double focal_x_baseline;
    int32_t setup_type; /* 0 Monocular, 1 Stereo, 2 RGBD */
} orc_pose_cam;

typedef struct orc_pt_obs { /* one matched keypoint (pose_optimizer.cc:126-151) */
    double pos_w[3];        /* lm->get_pos_in_world() */
    float obs_x, obs_y;     /* undist_keypts_[idx].pt */
    float x_right;          /* stereo_x_right_[idx]; < 0 => monocular edge */
    float inv_sigma_sq;     /* inv_level_sigma_sq_[octave] */
} orc_pt_obs;

typedef struct orc_line_obs { /* one matched keyline (pose_optimizer_extended_line.cc:160-188) */
    double plucker[6];        /* Line::get_PlueckerCoord(): (n, d) */
    float sp_x, sp_y, ep_x, ep_y;
    float inv_sigma_sq;
    float pad;
} orc_line_obs;

/* optimize::pose_optimizer::optimize / pose_optimizer_extended_line::optimize.  Returns the number of inlier
 * point observations; T_cw_out, pt_outlier[n_pts], line_outlier[n_lines] as written to the frame. */
int orc_pose_optimize(const orc_pose_cam *cam, const double *T_cw_in, const orc_pt_obs *pts, int n_pts,
                      const orc_line_obs *lines, int n_lines, int num_trials, int num_each_iter, double *T_cw_out,
                      uint8_t *pt_outlier, uint8_t *line_outlier, int *lm_iterations_out);

#ifdef __cplusplus
}
#endif
#endif
