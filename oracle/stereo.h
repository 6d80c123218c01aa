/* oracle/stereo.h -- match::stereo restatement (TEST INFRASTRUCTURE ONLY); see oracle.h for the rules. */
#ifndef PLP_ORACLE_STEREO_H
#define PLP_ORACLE_STEREO_H
#include <stdint.h>
#include "orb.h"
#ifdef __cplusplus
extern "C" {
#endif
/* match/stereo.cc:45-150.  The two pyramids are the concatenated levels (tight pitch) orc_orb_extract returns.
 * best_right_out (optional): index of the Hamming-closest right keypoint before the sub-pixel stage, -1 if none. */
void orc_stereo_compute(const uint8_t *pyr_left, const uint8_t *pyr_right, const int32_t *lvl_w, const int32_t *lvl_h,
                        int num_levels, const orc_keypoint *kp_l, const uint8_t *desc_l, int n_l, const orc_keypoint *kp_r,
                        const uint8_t *desc_r, int n_r, const float *scale_factors, const float *inv_scale_factors,
                        float focal_x_baseline, float true_baseline, float *x_right_out, float *depth_out,
                        int32_t *best_right_out);
#ifdef __cplusplus
}
#endif
#endif
