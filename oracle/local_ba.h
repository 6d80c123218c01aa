/* oracle/local_ba.h -- local bundle adjustment restatement (TEST INFRASTRUCTURE ONLY). */
#ifndef PLP_ORACLE_LOCAL_BA_H
#define PLP_ORACLE_LOCAL_BA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* The gathered graph of local_bundle_adjuster.cc:72-272 (keyframes = local + fixed, local landmarks, one edge per
 * observation).  Same layout as plp_ba_problem in include/plpslam_b200.h. */
typedef struct orc_ba_problem {
    double fx, fy, cx, cy, focal_x_baseline;
    int32_t setup_type; /* 0 Monocular (Huber delta sqrt(5.991)), else sqrt(7.815) for point edges */
    int32_t n_kf;
    const double *kf_pose_cw; /* n_kf x 16 */
    const uint8_t *kf_fixed;  /* id == 0 or "fixed keyframe" */
    int32_t n_pts;
    const double *pt_pos_w; /* n_pts x 3 */
    int32_t n_pt_edges;     /* grouped by landmark, in the order of local_bundle_adjuster.cc:226-272 */
    const int32_t *pt_edge_kf, *pt_edge_lm;
    const float *pt_edge_obs; /* x 3: x, y, x_right (< 0: monocular edge) */
    const float *pt_edge_inv_sigma_sq;
    int32_t n_lines;
    const double *line_plucker; /* n_lines x 6 */
    int32_t n_line_edges;
    const int32_t *line_edge_kf, *line_edge_lm;
    const float *line_edge_obs; /* x 4: sp.x, sp.y, ep.x, ep.y */
    const float *line_edge_inv_sigma_sq;
    int32_t n_plane_edges; /* point_plane_distance_edge: unary on a point landmark */
    const int32_t *plane_edge_lm;
    const double *plane_edge_fn; /* x 4: (n, d) */
} orc_ba_problem;

typedef struct orc_ba_result {
    double *kf_pose_cw;       /* n_kf x 16 */
    double *pt_pos_w;         /* n_pts x 3 */
    double *line_plucker;     /* n_lines x 6 */
    uint8_t *pt_edge_outlier; /* outlier_observations (local_bundle_adjuster.cc:342-372) */
    uint8_t *line_edge_outlier;
    int32_t iters_first, iters_second, lm_tries;
    double final_chi2;
} orc_ba_result;

int orc_local_ba(const orc_ba_problem *p, int num_first_iter, int num_second_iter, const volatile uint8_t *force_stop,
                 orc_ba_result *r);

/* optimize/global_bundle_adjuster.cc:64-253 on the same problem layout (kf_fixed = keyframe id 0 only) */
int orc_global_ba(const orc_ba_problem *p, int num_iter, int use_huber_kernel, const volatile uint8_t *force_stop,
                  orc_ba_result *r);

#ifdef __cplusplus
}
#endif
#endif
