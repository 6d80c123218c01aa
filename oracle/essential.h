/* oracle/essential.h -- solve::essential_solver::find_via_ransac restatement (TEST INFRASTRUCTURE ONLY); see oracle.h. */
#ifndef PLP_ORACLE_ESSENTIAL_H
#define PLP_ORACLE_ESSENTIAL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* solve/essential_solver.cc:123-160 compute_E_21 (n >= 8 correspondences, bearings n x 3) */
void orc_essential_compute_E21(const double *bearings_1, const double *bearings_2, int n, double *E_21_out);
/* solve/essential_solver.cc:37-121 find_via_ransac(max_num_iter = num_iter, recompute).  `samples` holds the
 * num_iter x 8 match indices of util::create_random_array(8, 0, num_matches - 1) (the reference seeds a fresh mt19937 from
 * std::random_device per call, util/random_array.cc:37-44, so its own result is not reproducible).  Returns
 * solution_is_valid; scores_out (optional) = the num_iter hypothesis scores. */
int orc_essential_ransac(const double *bearings_1, const double *bearings_2, const int32_t *matches_12, int num_matches,
                         const int32_t *samples, int num_iter, int recompute, uint8_t *is_inlier_out,
                         double *best_E_21_out, double *best_score_out, float *scores_out);
#ifdef __cplusplus
}
#endif
#endif
