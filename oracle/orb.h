/* oracle/orb.h -- ORB extractor restatement (TEST INFRASTRUCTURE ONLY); see oracle.h for the rules. */
#ifndef PLP_ORACLE_ORB_H
#define PLP_ORACLE_ORB_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_keypoint { /* cv::KeyPoint layout, 28 bytes */
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orc_keypoint;

typedef struct orc_orb_params { /* feature/orb_params.h */
    uint32_t max_num_keypts;
    float scale_factor;
    uint32_t num_levels, ini_fast_thr, min_fast_thr;
} orc_orb_params;

/* OpenCV primitives restated (pinned bit-exactly against cv2 in tests/test_orb_oracle.py) */
void orc_resize_linear(const uint8_t *src, int sw, int sh, int sstep, uint8_t *dst, int dw, int dh, int dstep);
int orc_fast9_16(const uint8_t *img, int w, int h, int step, int thr, int nonmax, orc_keypoint *out, int cap);
void orc_gaussian_blur_7x7(const uint8_t *src, int w, int h, int sstep, uint8_t *dst, int dstep);
void orc_gaussian_blur_5x5(const uint8_t *src, int w, int h, int sstep, uint8_t *dst, int dstep);
float orc_fast_atan2(float y, float x);
/* util/trigonometric.h:42-78 */
float orc_util_cos(float v);
float orc_util_sin(float v);

/* orb_params.cc:86-128, orb_extractor.cc:235-287 */
void orc_orb_tables(const orc_orb_params *p, float *scale_factors, float *inv_scale_factors, float *level_sigma_sq,
                    float *inv_level_sigma_sq, uint32_t *num_keypts_per_level, int32_t *u_max16);
void orc_orb_level_sizes(const orc_orb_params *p, int rows, int cols, int32_t *w_out, int32_t *h_out);
/* orb_extractor.cc:468-685 */
int orc_orb_distribute(const orc_orb_params *p, const orc_keypoint *cands, int n, int min_x, int max_x, int min_y,
                       int max_y, unsigned num_keypts, orc_keypoint *out);
/* orb_extractor.cc:708-735 / 747-807 */
float orc_orb_ic_angle(const orc_orb_params *p, const uint8_t *img, int w, int h, float px, float py);
void orc_orb_describe(const orc_orb_params *p, const uint8_t *blurred, int w, int h, const orc_keypoint *kp,
                      uint8_t *desc);
/* orb_extractor.cc:73-160.  Returns the number of keypoints (-1 if cap is too small).  Optional debug
 * outputs: concatenated pyramid levels, per-level FAST candidates (level-relative to the 19-px border,
 * in cell-row/cell-col/row-major order) and counts. */
int orc_orb_extract(const orc_orb_params *p, const uint8_t *img, int rows, int cols, int step, const uint8_t *mask,
                    int mask_step, orc_keypoint *kps_out, uint8_t *desc_out, int cap, uint8_t *pyramid_out,
                    orc_keypoint *cands_out, int cands_cap, int32_t *cands_per_level, int32_t *kps_per_level);

#ifdef __cplusplus
}
#endif
#endif
