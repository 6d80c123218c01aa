/* oracle/plane.h -- Planar_Mapping_module plane RANSAC restatement (TEST INFRASTRUCTURE ONLY); see oracle.h. */
#ifndef PLP_ORACLE_PLANE_H
#define PLP_ORACLE_PLANE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orc_plane_cfg {
    int32_t mode;              /* 0: estimate_plane_sequential_RANSAC (:412-591), 1: update_plane_via_RANSAC (:593-733) */
    int32_t points_per_ransac; /* POINTS_PER_RANSAC */
    double planar_distance_thresh, final_error_thresh, inliers_ratio_thr;
    double initial_best_error; /* mode 1: plane->get_best_error() */
} orc_plane_cfg;
/* planar_mapping_module.cc:735-771 over an index list; returns the residual */
double orc_plane_fit(const double *pos_w, const int32_t *idx, int cnt, double *eq_out);
/* The two RANSAC loops with the random index draws as an input (num_iter x sample_size indices; the reference draws them
 * from a std::random_device-seeded mt19937).  valid[j] = !lms[j]->will_be_erased().  eq_out / plane_error_out = the plane
 * equation and best_error_ the Plane object holds after the call (it is mutated every iteration, also on failure);
 * inlier_out[j] = 1 for the landmarks step [4] keeps.  Returns 1 (true), 0 (false), 2 (false + set_invalid). */
int orc_plane_ransac(const double *pos_w, const uint8_t *valid, int n, const int32_t *samples, int num_iter, int sample_size,
                     const orc_plane_cfg *cfg, double *eq_out, double *plane_error_out, uint8_t *inlier_out);
#ifdef __cplusplus
}
#endif
#endif
