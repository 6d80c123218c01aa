/* oracle/lines.h -- CPU restatement of the LSD + LBD line front end (TEST INFRASTRUCTURE ONLY, see oracle.h).
 *
 * Follows feature/line_extractor.cc:88-160, feature/line_descriptor/LSDDetector_custom.cpp:216-320 and
 * feature/line_descriptor/binary_descriptor_custom.cpp:217-258, 347-408, 518-679, 1018-1364 of the reference, plus the
 * third-party cv::LineSegmentDetector (OpenCV imgproc lsd.cpp, NOT part of /root/reference) restated from its
 * published algorithm (Grompone von Gioi et al., "LSD: a Line Segment Detector", IPOL 2012) and pinned bit-exactly
 * against cv2 4.13 `createLineSegmentDetector(1, 0.5, 0.6, 2.0, 22.5, 1.0, 0.6, 1024).detect` in
 * tests/test_lines_oracle.py.
 */
#ifndef PLP_ORACLE_LINES_H
#define PLP_ORACLE_LINES_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_keyline { /* cv::line_descriptor::KeyLine, descriptor_custom.hpp:105-199 (68 bytes) */
    float angle;
    int32_t class_id;
    int32_t octave;
    float pt_x, pt_y;
    float response;
    float size;
    float start_x, start_y, end_x, end_y;
    float s_oct_x, s_oct_y, e_oct_x, e_oct_y;
    float line_length;
    int32_t num_pixels;
} orc_keyline;

/* seed_order: 0 = pseudo-ordering by gradient bin with raster order inside a bin (the bin lists of the original LSD and
 *                 of OpenCV 3.4, the version the reference's README names) -- the order the CUDA path implements;
 *             1 = std::sort on the bin only (OpenCV >= 4.5: unstable, order inside a bin is whatever libstdc++'s
 *                 introsort leaves) -- used to pin this restatement against cv2 4.13.
 * libm_float: 1 = cosf/sinf of glibc where cv calls cos(float)/sin(float); 0 = evaluate in double, round to float
 *                 (the determinism rule the CUDA path can follow). */
/* sum_order:  0 = sequential double sums and swap-remove compaction as lsd.cpp writes them; 1 = 32 strided partial sums
 *                 combined by an xor tree and order-preserving compaction (what a warp computes) -- the CUDA path. */
typedef struct orc_lsd_config {
    int32_t seed_order;
    int32_t libm_float;
    int32_t sum_order;
} orc_lsd_config;

/* GaussianBlur(11x11, sigma 1.2) fixed-point + resize(0.5, INTER_LINEAR_EXACT): the image LSD works on */
void orc_lsd_scaled_image(const uint8_t *img, int w, int h, int step, uint8_t *out /* (w/2) x (h/2) */);
/* level-line angle (degrees from cv::fastAtan2, -1024 = NOTDEF), 4*modgrad^2 (= gx^2+gy^2) and the seed order;
 * returns the number of seeds (all pixels of the (w-1) x (h-1) interior, as cv orders them) */
int orc_lsd_ll_angle(const uint8_t *scaled, int w, int h, const orc_lsd_config *cfg, float *angle_deg, int32_t *grad_sq,
                     int32_t *bins, int32_t *order);
/* full detector on the full-resolution image: segments as (x1,y1,x2,y2) float, in detection order */
int orc_lsd_detect(const uint8_t *img, int w, int h, int step, const orc_lsd_config *cfg, float *segments, int cap);
/* debug: per processed seed {order position, x, y, first region size, pixels accepted in total, final region size, bbox x0 y0
 * x1 y1 of everything accepted, segment emitted}; returns the number of records (-needed if cap is too small) */
int orc_debug_lsd_trace(const uint8_t *img, int w, int h, int step, const orc_lsd_config *cfg, int32_t *out, int cap_records);
/* LSDDetectorC::detectImpl (LSDDetector_custom.cpp:225-320) for one octave */
int orc_lsd_keylines(const uint8_t *img, int w, int h, int step, const orc_lsd_config *cfg, double min_length,
                     orc_keyline *out, int cap);
/* cv::Sobel(3x3, CV_16S) of the GaussianBlur(5x5, 1) image (binary_descriptor_custom.cpp:347-395) */
void orc_lbd_gradients(const uint8_t *img, int w, int h, int step, int16_t *dx, int16_t *dy);
/* BinaryDescriptor::compute (binary_descriptor_custom.cpp:518-679, 1018-1364); desc_float (72 per line) optional */
void orc_lbd_compute(const uint8_t *img, int w, int h, int step, const orc_keyline *kl, int n, int libm_float,
                     uint8_t *desc, float *desc_float);
/* LineFeatureTracker::extract_LSD_LBD (line_extractor.cc:88-160): keylines with octave 0 and length >= 60, their LBD
 * rows and 2-D line functions; returns the count (or -needed if cap is too small) */
int orc_line_extract(const uint8_t *img, int w, int h, int step, const orc_lsd_config *cfg, orc_keyline *kl_out,
                     uint8_t *lbd_out, double *fn_out, int cap);

#ifdef __cplusplus
}
#endif
#endif
