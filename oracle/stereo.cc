// oracle/stereo.cc -- CPU restatement of match::stereo::compute (TEST INFRASTRUCTURE ONLY, see oracle.h).
//
// Follows /root/reference/src/PLPSLAM/match/stereo.cc:45-302 line by line (OpenMP off, the reference default).
// cv::norm(L1) of the centre-subtracted float patches is a sum of integer-valued floats below 2^24, hence exact in any
// order; it is restated as an integer sum.
#include "stereo.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <utility>
#include <vector>

#include "oracle.h"

namespace {
inline int cv_floor(double v) {
    int i = (int)v;
    return i - (i > v);
}
inline int cv_ceil(double v) {
    int i = (int)v;
    return i + (i < v);
}
struct Pyr {
    const uint8_t *base;
    const int32_t *w, *h;
    std::vector<size_t> off;
    const uint8_t *px(int l, int y, int x) const { return base + off[l] + (size_t)y * w[l] + x; }
};
constexpr unsigned kHammThr = (100 + 50) / 2;  // stereo.h:126
}  // namespace

extern "C" void orc_stereo_compute(const uint8_t *pyr_left, const uint8_t *pyr_right, const int32_t *lvl_w,
                                   const int32_t *lvl_h, int num_levels, const orc_keypoint *kp_l, const uint8_t *desc_l,
                                   int n_l, const orc_keypoint *kp_r, const uint8_t *desc_r, int n_r,
                                   const float *scale_factors, const float *inv_scale_factors, float focal_x_baseline,
                                   float true_baseline, float *x_right_out, float *depth_out, int32_t *best_right_out) {
    Pyr L{pyr_left, lvl_w, lvl_h, {}}, R{pyr_right, lvl_w, lvl_h, {}};
    size_t o = 0;
    for (int l = 0; l < num_levels; ++l) {
        L.off.push_back(o);
        R.off.push_back(o);
        o += (size_t)lvl_w[l] * lvl_h[l];
    }
    const float min_disp = 0.0f, max_disp = focal_x_baseline / true_baseline;  // stereo.cc:42
    // get_right_keypoint_indices_in_each_row(2.0), stereo.cc:152-183
    const int rows = lvl_h[0];
    std::vector<std::vector<unsigned>> in_row(rows);
    for (int ir = 0; ir < n_r; ++ir) {
        const float y = kp_r[ir].y;
        const float r = 2.0f * scale_factors[kp_r[ir].octave];
        const int max_r = cv_ceil(y + r), min_r = cv_floor(y - r);
        for (int row = min_r; row <= max_r; ++row)
            if (row >= 0 && row < rows) in_row[row].push_back((unsigned)ir);  // .at() would throw outside; never happens
    }
    for (int i = 0; i < n_l; ++i) {
        x_right_out[i] = -1.0f;
        depth_out[i] = -1.0f;
        if (best_right_out) best_right_out[i] = -1;
    }
    std::vector<std::pair<int, int>> corr_idx;
    for (int il = 0; il < n_l; ++il) {
        const orc_keypoint &kl = kp_l[il];
        const int lvl = kl.octave;
        const float y_left = kl.y, x_left = kl.x;
        const auto &cands = in_row.at((size_t)y_left);
        if (cands.empty()) continue;
        const float min_x_right = x_left - max_disp, max_x_right = x_left - min_disp;
        if (max_x_right < 0) continue;
        // find_closest_keypoints_in_stereo, stereo.cc:185-224
        unsigned best_ir = 0, best_d = kHammThr;
        for (unsigned ir : cands) {
            const orc_keypoint &kr = kp_r[ir];
            if (kr.octave < lvl - 1 || kr.octave > lvl + 1) continue;
            if (kr.x < min_x_right || max_x_right < kr.x) continue;
            const unsigned d = orc_hamming_32(desc_l + (size_t)il * 32, desc_r + (size_t)ir * 32);
            if (d < best_d) {
                best_ir = ir;
                best_d = d;
            }
        }
        if (kHammThr <= best_d) continue;
        if (best_right_out) best_right_out[il] = (int)best_ir;
        // compute_subpixel_disparity, stereo.cc:226-299
        const float x_right = kp_r[best_ir].x;
        const float isf = inv_scale_factors[lvl];
        const int sxl = (int)std::lrintf(kl.x * isf), syl = (int)std::lrintf(kl.y * isf), sxr = (int)std::lrintf(x_right * isf);
        constexpr int win = 5, slide = 5;
        const int ini_x = sxr - slide - win, end_x = sxr + slide + win;
        if (ini_x < 0 || lvl_w[lvl] <= end_x) continue;
        float best_corr = (float)UINT_MAX;
        int best_off = 0;
        float corr[2 * slide + 1];
        const int lc = *L.px(lvl, syl, sxl);
        for (int off = -slide; off <= slide; ++off) {
            const int rc = *R.px(lvl, syl, sxr + off);
            long sum = 0;
            for (int dy = -win; dy <= win; ++dy)
                for (int dx = -win; dx <= win; ++dx)
                    sum += std::abs((*L.px(lvl, syl + dy, sxl + dx) - lc) - (*R.px(lvl, syl + dy, sxr + off + dx) - rc));
            const float c = (float)sum;
            if (c < best_corr) {
                best_corr = c;
                best_off = off;
            }
            corr[slide + off] = c;
        }
        if (best_off == -slide || best_off == slide) continue;
        const float c1 = corr[slide + best_off - 1], c2 = corr[slide + best_off], c3 = corr[slide + best_off + 1];
        const float x_delta = (float)((c1 - c3) / (2.0 * (c1 + c3) - 4.0 * c2));
        if (x_delta < -1.0 || 1.0 < x_delta) continue;
        float best_x_right = scale_factors[lvl] * ((float)(sxr + best_off) + x_delta);
        float best_disp = kl.x - best_x_right;
        if (best_disp < min_disp || max_disp <= best_disp) continue;
        if (best_disp <= 0.0f) {
            best_disp = 0.01f;
            best_x_right = x_left - best_disp;
        }
        depth_out[il] = focal_x_baseline / best_disp;
        x_right_out[il] = best_x_right;
        corr_idx.emplace_back((int)best_corr, il);
    }
    // stereo.cc:124-148: reject correlations above twice the median
    std::sort(corr_idx.begin(), corr_idx.end());
    const size_t median_i = corr_idx.size() / 2;
    const float median = corr_idx.empty() ? 0.0f : (float)corr_idx[median_i].first;
    const float thr = (float)(2.0 * median);
    for (size_t i = median_i; i < corr_idx.size(); ++i)
        if (thr < (float)corr_idx[i].first) {
            x_right_out[corr_idx[i].second] = -1;
            depth_out[corr_idx[i].second] = -1;
        }
}
