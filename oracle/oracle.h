/*
 * oracle.h -- CPU restatement of the Structure-PLP-SLAM hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This library is the parity checker for libplpslam_b200.so.  It is a plain, sequential
 * C++17 restatement of the reference algorithms, each function citing the reference file:line
 * it follows (paths relative to /root/reference/src/PLPSLAM).  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it; the product
 * (structure-plp-slam_b200/) never does.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - Hamming, scale tables, angle checker, cell indices, trig: pinned by the reference's own
 *     known-answer tests (test/PLPSLAM/match/base.cc etc.), restated in tests/.
 *   - OpenCV primitives (resize, FAST, GaussianBlur, fastAtan2, Sobel): pinned bit-exactly
 *     against cv2 4.13 (the third-party library the reference calls) in tests/.
 *   - g2o Levenberg-Marquardt / Schur / Huber semantics: PARITY UNPINNED -- no g2o source or
 *     binary exists in this environment; restated from the published algorithm and anchored on
 *     the reference's call sites only.
 *
 * Determinism rules where the reference is under-determined: IEEE-754 without FMA contraction
 * (-ffp-contract=off), cvRound = round-half-even, quadtree ties broken by node creation order,
 * angle-histogram ties by (size desc, bin asc), unordered_map iteration replaced by ascending id.
 */
#ifndef PLP_ORACLE_H
#define PLP_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#include "orb.h"
#include "pose_opt.h"
#include "local_ba.h"
#include "lines.h"
#include "stereo.h"
#include "bow.h"
#include "essential.h"
#include "plane.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- match/base.h:43-93 ------------------------------------------------------------ */
unsigned orc_hamming_32(const uint8_t *a, const uint8_t *b);
unsigned orc_hamming_64(const uint8_t *a, const uint8_t *b);
void orc_hamming_matrix(const uint8_t *a, int na, const uint8_t *b, int nb, uint16_t *out);
void orc_hamming_nn(const uint8_t *q, int nq, const uint8_t *t, int nt, int32_t *idx, uint16_t *dist);

/* ---- match/angle_checker.h:86-175 -------------------------------------------------- */
/* returns number of invalid matches written to invalid_out (capacity n) */
int orc_angle_checker_invalid(const float *delta_angles, const int32_t *matches, int n,
                              int histogram_length, int num_bins_thr, int32_t *invalid_out);
int orc_angle_checker_valid(const float *delta_angles, const int32_t *matches, int n,
                            int histogram_length, int num_bins_thr, int32_t *valid_out);

/* ---- data/common.h:104-109, data/common.cc:205-313 ----------------------------------- */
typedef struct orc_grid {
    float min_x, min_y;
    double inv_cell_width, inv_cell_height;
    int32_t num_cols, num_rows;
} orc_grid;
int orc_get_cell_indices(const orc_grid *g, float x, float y, int *cx, int *cy);
/* builds the grid over (x,y), then returns the indices in traversal order; returns count */
int orc_get_keypoints_in_cell(const orc_grid *g, const float *x, const float *y,
                              const int32_t *octave, int n, float ref_x, float ref_y, float margin,
                              int min_level, int max_level, int32_t *indices_out);

typedef struct orc_camera {
    double fx, fy, cx, cy;
    double focal_x_baseline;
    double true_baseline;
    float min_x, max_x, min_y, max_y;
    int32_t setup_type;
} orc_camera;
/* camera/perspective.cc:190-209 */
int orc_reproject_to_image(const orc_camera *cam, const double *rot_cw /*3x3 row-major*/,
                           const double *trans_cw, const double *pos_w, double *reproj /*2*/,
                           float *x_right);

/* ---- match/projection.cc:37-121 ------------------------------------------------------ */
unsigned orc_match_frame_and_landmarks(const orc_grid *g, int n, const float *x, const float *y,
                                       const int32_t *octave, const float *x_right,
                                       const uint8_t *desc, const uint8_t *claimed,
                                       const float *scale_factors, int num_levels, int m,
                                       const float *reproj_x, const float *reproj_y,
                                       const float *q_x_right, const int32_t *scale_level,
                                       const uint8_t *q_desc, const uint8_t *q_valid, float margin,
                                       float lowe_ratio, int32_t *best_idx_out);

/* ---- match/projection.cc:214-358 ----------------------------------------------------- */
unsigned orc_match_current_and_last_frames(
    const orc_grid *g, int n, const float *x, const float *y, const int32_t *octave,
    const float *angle, const float *x_right, const uint8_t *desc, const uint8_t *claimed,
    const float *scale_factors, int num_levels, const orc_camera *cam, const double *pose_cw_curr,
    const double *pose_cw_last, int n_last, const double *pos_w, const int32_t *last_octave,
    const float *last_angle, const uint8_t *last_desc, const uint8_t *last_valid, float margin,
    int check_orientation, int32_t *matched_last_idx_out);

/* ---- match/projection.cc:124-212 ----------------------------------------------------- */
unsigned orc_match_frame_and_landmarks_line(int n, const float *sx, const float *sy,
                                            const float *ex, const float *ey,
                                            const int32_t *octave, const int32_t *ratio_level,
                                            const uint8_t *desc, const uint8_t *claimed,
                                            const float *scale_factors_lsd, int num_levels_lsd,
                                            int m, const float *sp_x, const float *sp_y,
                                            const float *ep_x, const float *ep_y,
                                            const int32_t *scale_level, const uint8_t *q_desc,
                                            const uint8_t *q_valid, float margin, float lowe_ratio,
                                            int32_t *best_idx_out);

/* ---- match/projection.cc:361-527 ----------------------------------------------------- */
unsigned orc_match_current_and_last_frames_line(
    int n, const float *sx, const float *sy, const float *ex, const float *ey,
    const int32_t *octave, const float *x_right_sp, const float *x_right_ep, const uint8_t *desc,
    const uint8_t *claimed, const float *scale_factors_lsd, int num_levels_lsd,
    const orc_camera *cam, const double *pose_cw_curr, const double *pose_cw_last, int n_last,
    const double *pos_w /*n_last x 6*/, const int32_t *last_octave, const uint8_t *last_desc,
    const uint8_t *last_valid, float margin, int32_t *matched_last_idx_out);

/* ---- match/projection.cc:529-645 / 648-779 (relocalisation matchers).  kf_valid[idx] = lm && !will_be_erased &&
 * !already_matched.  The optional q_* outputs are the flattened per-landmark queries the reference-side adapter would
 * hand to the C ABI (reprojection, predicted scale level, survived the visibility / distance gates). */
unsigned orc_match_frame_and_keyframe(const orc_grid *g, int n, const float *x, const float *y, const int32_t *octave,
                                      const float *angle, const uint8_t *desc, const uint8_t *claimed,
                                      const float *scale_factors, int num_levels, float log_scale_factor,
                                      const orc_camera *cam, const double *pose_cw_curr, int n_kf, const double *pos_w,
                                      const float *min_valid_dist, const float *max_valid_dist, const float *kf_angle,
                                      const uint8_t *kf_desc, const uint8_t *kf_valid, float margin, unsigned hamm_dist_thr,
                                      int check_orientation, int32_t *matched_kf_idx_out, float *q_reproj_x,
                                      float *q_reproj_y, int32_t *q_level, uint8_t *q_valid);
unsigned orc_match_frame_and_keyframe_line(int n, const float *sx, const float *sy, const float *ex, const float *ey,
                                           const int32_t *octave, const uint8_t *desc, const uint8_t *claimed,
                                           const float *scale_factors_lsd, int num_levels_lsd, float log_scale_factor_lsd,
                                           const orc_camera *cam, const double *pose_cw_curr, int n_kf,
                                           const double *pos_w, const float *min_valid_dist, const float *max_valid_dist,
                                           const uint8_t *kf_desc, const uint8_t *kf_valid, float margin,
                                           unsigned hamm_dist_thr, int32_t *matched_kf_idx_out, float *q_sp_x,
                                           float *q_sp_y, float *q_ep_x, float *q_ep_y, int32_t *q_level, uint8_t *q_valid);

/* ---- match/robust.cc:43-216 (+ 387-406).  Feature vectors flattened in iteration order (ascending node id). */
unsigned orc_match_for_triangulation(int n1, const uint8_t *desc1, const float *angle1, const int32_t *octave1,
                                     const double *bearing1, const uint8_t *has_lm1, const float *x_right1, int n2,
                                     const uint8_t *desc2, const float *angle2, const double *bearing2,
                                     const uint8_t *has_lm2, const float *x_right2, int nodes1, const uint32_t *ids1,
                                     const int32_t *off1, const uint32_t *idx1, int nodes2, const uint32_t *ids2,
                                     const int32_t *off2, const uint32_t *idx2, const double *E_12, const double *epipole,
                                     const float *scale_factors_1, int check_orientation, int libm,
                                     int32_t *matched_idx2_in_1_out);

/* ---- data/landmark.cc:181-247 (and data/landmark_line.cc:215-283) */
void orc_landmark_compute_descriptor_batch(const uint8_t *descs, const int32_t *offsets, int num_landmarks,
                                           int32_t *best_idx_out);

/* ---- match/fuse.cc:40-151 (mode 0, detect_duplication) / :153-300 (mode 1, replace_duplication): the per-landmark search
 * of ONE target keyframe (everything up to "auto *lm_in_keyfrm = keyfrm->get_landmark(best_idx)"); best_idx_out[i] = -1
 * wherever the reference loop body `continue`s.  The sequential effects (add_observation / replace) belong to the caller.
 * level_out (optional): predicted scale level of the landmarks that reached predict_scale_level, else -1. */
void orc_fuse_search_points(const orc_grid *g, const orc_camera *cam, int n, const float *x, const float *y,
                            const int32_t *octave, const float *x_right, const uint8_t *desc, const double *rot_cw,
                            const double *trans_cw, const double *cam_center, const float *scale_factors,
                            const float *inv_level_sigma_sq, int num_levels, float log_scale_factor, int m,
                            const double *pos_w, const double *obs_mean_normal, const float *min_valid_dist,
                            const float *max_valid_dist, const float *max_valid_dist_raw, const uint8_t *lm_desc,
                            const uint8_t *lm_valid, const uint8_t *lm_skip, float margin, int mode,
                            int32_t *best_idx_out, uint16_t *best_dist_out, int32_t *level_out);
/* ---- match/fuse.cc:304-503 (replace_duplication_line), same contract */
void orc_fuse_search_lines(const orc_camera *cam, int n, const float *sx, const float *sy, const float *ex, const float *ey,
                           const int32_t *octave, const uint8_t *desc, const double *rot_cw, const double *trans_cw,
                           const double *cam_center, const float *scale_factors_lsd, const float *inv_level_sigma_sq_lsd,
                           int num_levels_lsd, float log_scale_factor_lsd, int m, const double *pos_w /*m x 6*/,
                           const float *min_valid_dist, const float *max_valid_dist, const float *max_valid_dist_raw,
                           const uint8_t *lm_desc, const uint8_t *lm_valid, const uint8_t *lm_skip, float margin,
                           int32_t *best_idx_out, uint16_t *best_dist_out, int32_t *level_out);
/* data/landmark.cc:319-362 predict_scale_level (host libm logf) -- exposed for the threshold-table test */
unsigned orc_predict_scale_level(float max_valid_dist, float cam_to_lm_dist, float log_scale_factor, unsigned num_levels);

/* ---- match/bow_tree.cc:41-165 (match_frame_and_keyframe: side 1 = keyframe, side 2 = frame, valid2 = NULL) and
 * :167-305 (match_keyframes: side 1 = keyfrm_1, side 2 = keyfrm_2, valid2 = lm_2 && !will_be_erased).  valid1[i] =
 * lm_1 && !will_be_erased.  Feature vectors flattened in iteration order (ascending node id).  Outputs after the
 * orientation check: matched_2_of_1[n1] (index on side 2 matched to each side-1 keypoint) and matched_1_of_2[n2]. */
unsigned orc_bow_tree_match(int n1, const uint8_t *desc1, const float *angle1, const uint8_t *valid1, int n2,
                            const uint8_t *desc2, const float *angle2, const uint8_t *valid2, int nodes1,
                            const uint32_t *ids1, const int32_t *off1, const uint32_t *idx1, int nodes2,
                            const uint32_t *ids2, const int32_t *off2, const uint32_t *idx2, float lowe_ratio,
                            int check_orientation, int32_t *matched_2_of_1, int32_t *matched_1_of_2);

/* ---- match/robust.cc:257-385 --------------------------------------------------------- */
unsigned orc_brute_force_match(const uint8_t *frm_desc, const float *frm_angle, int n_frm,
                               const uint8_t *kf_desc, const float *kf_angle,
                               const uint8_t *kf_valid, int n_kf, float lowe_ratio,
                               int check_orientation, int32_t *matched_kf_idx_in_frm_out);

#ifdef __cplusplus
}
#endif
#endif
