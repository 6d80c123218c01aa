// oracle/match.cc -- CPU restatement of the Hamming matchers (TEST INFRASTRUCTURE ONLY).
// Follows /root/reference/src/PLPSLAM/match/{base.h,angle_checker.h,projection.cc,robust.cc},
// data/common.{h,cc} and camera/perspective.cc; see oracle.h for the pinning status.
#include "oracle.h"
#include "detmath.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

namespace {

// OpenCV scalar helpers (cvFloor/cvCeil/cvRound; cvRound = round-half-even)
inline int cvFloor(double v) {
    int i = (int)v;
    return i - (i > v);
}
inline int cvCeil(double v) {
    int i = (int)v;
    return i + (i < v);
}
inline int cvRoundF(float v) { return (int)std::lrintf(v); }

constexpr unsigned HAMMING_DIST_THR_LOW = 50;    // match/base.h:38
constexpr unsigned HAMMING_DIST_THR_HIGH = 100;  // match/base.h:39
constexpr unsigned MAX_HAMMING_DIST = 256;       // match/base.h:40

// match/angle_checker.h:86-175 with the oracle's stable bin ranking
struct AngleChecker {
    unsigned histogram_length;
    float inv_histogram_length;
    unsigned num_bins_thr;
    std::vector<std::vector<int>> hist;
    AngleChecker(unsigned len = 30, unsigned thr = 3)
        : histogram_length(len), inv_histogram_length(1.0f / len), num_bins_thr(thr), hist(len) {}
    void append(float delta_angle, int match) {
        // angle_checker.h:100-113
        if (delta_angle < 0.0) delta_angle += 360.0;
        if (360.0 <= delta_angle) delta_angle -= 360.0;
        const auto bin = static_cast<unsigned>(cvRoundF(delta_angle * inv_histogram_length));
        hist.at(bin).push_back(match);
    }
    std::vector<unsigned> ranked_bins() const {
        // angle_checker.h:163-175 (std::sort by size desc; ties made stable: bin index asc)
        std::vector<unsigned> idx(hist.size());
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(),
                         [this](unsigned a, unsigned b) { return hist[a].size() > hist[b].size(); });
        return idx;
    }
    std::vector<int> collect(bool want_valid) const {
        // angle_checker.h:115-161
        std::vector<int> out;
        const auto bins = ranked_bins();
        for (unsigned bin = 0; bin < histogram_length; ++bin) {
            const bool is_valid =
                std::any_of(bins.begin(), bins.begin() + num_bins_thr, [bin](unsigned i) { return bin == i; });
            if (is_valid == want_valid) out.insert(out.end(), hist[bin].begin(), hist[bin].end());
        }
        return out;
    }
};

// data/common.cc:205-231
struct Grid {
    const orc_grid *g;
    std::vector<std::vector<std::vector<unsigned>>> cells;
    Grid(const orc_grid *g_, const float *x, const float *y, int n) : g(g_) {
        cells.resize(g->num_cols);
        for (auto &col : cells) col.resize(g->num_rows);
        for (int idx = 0; idx < n; ++idx) {
            int cx, cy;
            if (orc_get_cell_indices(g, x[idx], y[idx], &cx, &cy)) cells[cx][cy].push_back(idx);
        }
    }
    // data/common.cc:241-313
    std::vector<unsigned> query(const float *x, const float *y, const int32_t *octave, float ref_x,
                                float ref_y, float margin, int min_level, int max_level) const {
        std::vector<unsigned> indices;
        const int min_cell_idx_x = std::max(0, cvFloor((ref_x - g->min_x - margin) * g->inv_cell_width));
        if (g->num_cols <= min_cell_idx_x) return indices;
        const int max_cell_idx_x =
            std::min(g->num_cols - 1, cvCeil((ref_x - g->min_x + margin) * g->inv_cell_width));
        if (max_cell_idx_x < 0) return indices;
        const int min_cell_idx_y = std::max(0, cvFloor((ref_y - g->min_y - margin) * g->inv_cell_height));
        if (g->num_rows <= min_cell_idx_y) return indices;
        const int max_cell_idx_y =
            std::min(g->num_rows - 1, cvCeil((ref_y - g->min_y + margin) * g->inv_cell_height));
        if (max_cell_idx_y < 0) return indices;
        const bool check_level = (0 < min_level) || (0 <= max_level);
        for (int cx = min_cell_idx_x; cx <= max_cell_idx_x; ++cx) {
            for (int cy = min_cell_idx_y; cy <= max_cell_idx_y; ++cy) {
                for (unsigned idx : cells[cx][cy]) {
                    if (check_level) {
                        if (octave[idx] < min_level) continue;
                        if (0 <= max_level && max_level < octave[idx]) continue;
                    }
                    const float dist_x = x[idx] - ref_x;
                    const float dist_y = y[idx] - ref_y;
                    if (std::abs(dist_x) < margin && std::abs(dist_y) < margin) indices.push_back(idx);
                }
            }
        }
        return indices;
    }
};

// data/common.cc:315-364
std::vector<unsigned> keylines_in_cell(int n, const float *sx, const float *sy, const float *ex,
                                       const float *ey, const int32_t *octave, float ref_x1,
                                       float ref_y1, float ref_x2, float ref_y2, float margin,
                                       int min_level, int max_level) {
    std::vector<unsigned> indices;
    // Vec3_t point_sp{ref_x1, ref_y1, 1.0}.cross(point_ep) in double
    const double ax = ref_x1, ay = ref_y1, bx = ref_x2, by = ref_y2;
    const double l0 = ay * 1.0 - 1.0 * by;
    const double l1 = 1.0 * bx - ax * 1.0;
    const double l2 = ax * by - ay * bx;
    const bool check_level = (0 < min_level) || (0 <= max_level);
    for (int i = 0; i < n; ++i) {
        const double den = std::sqrt(l0 * l0 + l1 * l1);
        const float distance_sp = (float)((sx[i] * l0 + sy[i] * l1 + l2) / den);
        const float distance_ep = (float)((ex[i] * l0 + ey[i] * l1 + l2) / den);
        if (std::abs(distance_sp) > margin || std::abs(distance_ep) > margin) continue;
        if (check_level) {
            if (octave[i] < min_level) continue;
            if (max_level > 0 && octave[i] > max_level) continue;
        }
        indices.push_back(i);
    }
    return indices;
}

struct Pose {
    double R[9];
    double t[3];
    explicit Pose(const double *T) {
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) R[r * 3 + c] = T[r * 4 + c];
            t[r] = T[r * 4 + 3];
        }
    }
};

// projection.cc:220-238 : forward/backward assumption
void motion_assumption(const orc_camera *cam, const Pose &cw, const Pose &lw, bool *fwd, bool *bwd) {
    // trans_wc = -rot_cw^T * trans_cw ; trans_lc = rot_lw * trans_wc + trans_lw
    double twc[3];
    for (int r = 0; r < 3; ++r)
        twc[r] = -(cw.R[0 * 3 + r] * cw.t[0] + cw.R[1 * 3 + r] * cw.t[1] + cw.R[2 * 3 + r] * cw.t[2]);
    const double tlc_z = lw.R[6] * twc[0] + lw.R[7] * twc[1] + lw.R[8] * twc[2] + lw.t[2];
    const bool mono = cam->setup_type == 0;
    *fwd = mono ? false : tlc_z > cam->true_baseline;
    *bwd = mono ? false : -tlc_z > cam->true_baseline;
}

}  // namespace

extern "C" {

unsigned orc_hamming_32(const uint8_t *a, const uint8_t *b) {
    // match/base.h:43-67 (SWAR popcount over 8 x u32)
    constexpr uint32_t mask_1 = 0x55555555U, mask_2 = 0x33333333U, mask_3 = 0x0F0F0F0FU, mask_4 = 0x01010101U;
    uint32_t pa[8], pb[8];
    std::memcpy(pa, a, 32);
    std::memcpy(pb, b, 32);
    unsigned dist = 0;
    for (unsigned i = 0; i < 8; ++i) {
        auto v = pa[i] ^ pb[i];
        v -= ((v >> 1) & mask_1);
        v = (v & mask_2) + ((v >> 2) & mask_2);
        dist += (((v + (v >> 4)) & mask_3) * mask_4) >> 24;
    }
    return dist;
}

unsigned orc_hamming_64(const uint8_t *a, const uint8_t *b) {
    // match/base.h:70-93
    constexpr uint64_t mask_1 = 0x5555555555555555UL, mask_2 = 0x3333333333333333UL,
                       mask_3 = 0x0F0F0F0F0F0F0F0FUL, mask_4 = 0x0101010101010101UL;
    uint64_t pa[4], pb[4];
    std::memcpy(pa, a, 32);
    std::memcpy(pb, b, 32);
    unsigned dist = 0;
    for (unsigned i = 0; i < 4; ++i) {
        auto v = pa[i] ^ pb[i];
        v -= (v >> 1) & mask_1;
        v = (v & mask_2) + ((v >> 2) & mask_2);
        dist += (unsigned)((((v + (v >> 4)) & mask_3) * mask_4) >> 56);
    }
    return dist;
}

void orc_hamming_matrix(const uint8_t *a, int na, const uint8_t *b, int nb, uint16_t *out) {
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) out[(size_t)i * nb + j] = (uint16_t)orc_hamming_32(a + 32 * i, b + 32 * j);
}

void orc_hamming_nn(const uint8_t *q, int nq, const uint8_t *t, int nt, int32_t *idx, uint16_t *dist) {
    // binary_descriptor_matcher.cpp:197-254 semantics: exact 1-NN (ties -> lowest train index)
    for (int i = 0; i < nq; ++i) {
        int best = -1;
        unsigned bd = 1u << 30;
        for (int j = 0; j < nt; ++j) {
            const unsigned d = orc_hamming_32(q + 32 * i, t + 32 * j);
            if (d < bd) {
                bd = d;
                best = j;
            }
        }
        idx[i] = best;
        dist[i] = (uint16_t)(best < 0 ? 0xFFFF : bd);
    }
}

int orc_angle_checker_invalid(const float *delta_angles, const int32_t *matches, int n, int histogram_length,
                              int num_bins_thr, int32_t *invalid_out) {
    AngleChecker ac(histogram_length, num_bins_thr);
    for (int i = 0; i < n; ++i) ac.append(delta_angles[i], matches[i]);
    const auto v = ac.collect(false);
    std::copy(v.begin(), v.end(), invalid_out);
    return (int)v.size();
}

int orc_angle_checker_valid(const float *delta_angles, const int32_t *matches, int n, int histogram_length,
                            int num_bins_thr, int32_t *valid_out) {
    AngleChecker ac(histogram_length, num_bins_thr);
    for (int i = 0; i < n; ++i) ac.append(delta_angles[i], matches[i]);
    const auto v = ac.collect(true);
    std::copy(v.begin(), v.end(), valid_out);
    return (int)v.size();
}

int orc_get_cell_indices(const orc_grid *g, float x, float y, int *cx, int *cy) {
    // data/common.h:104-109
    *cx = cvFloor((x - g->min_x) * g->inv_cell_width);
    *cy = cvFloor((y - g->min_y) * g->inv_cell_height);
    return (0 <= *cx && *cx < g->num_cols && 0 <= *cy && *cy < g->num_rows);
}

int orc_get_keypoints_in_cell(const orc_grid *g, const float *x, const float *y, const int32_t *octave, int n,
                              float ref_x, float ref_y, float margin, int min_level, int max_level,
                              int32_t *indices_out) {
    Grid grid(g, x, y, n);
    const auto v = grid.query(x, y, octave, ref_x, ref_y, margin, min_level, max_level);
    for (size_t i = 0; i < v.size(); ++i) indices_out[i] = (int32_t)v[i];
    return (int)v.size();
}

int orc_reproject_to_image(const orc_camera *cam, const double *R, const double *t, const double *pos_w,
                           double *reproj, float *x_right) {
    // camera/perspective.cc:190-209
    const double pc0 = R[0] * pos_w[0] + R[1] * pos_w[1] + R[2] * pos_w[2] + t[0];
    const double pc1 = R[3] * pos_w[0] + R[4] * pos_w[1] + R[5] * pos_w[2] + t[1];
    const double pc2 = R[6] * pos_w[0] + R[7] * pos_w[1] + R[8] * pos_w[2] + t[2];
    if (pc2 <= 0.0) return 0;
    const double z_inv = 1.0 / pc2;
    reproj[0] = cam->fx * pc0 * z_inv + cam->cx;
    reproj[1] = cam->fy * pc1 * z_inv + cam->cy;
    *x_right = (float)(reproj[0] - cam->focal_x_baseline * z_inv);
    return (cam->min_x < reproj[0] && reproj[0] < cam->max_x && cam->min_y < reproj[1] && reproj[1] < cam->max_y);
}

unsigned orc_match_frame_and_landmarks(const orc_grid *g, int n, const float *x, const float *y,
                                       const int32_t *octave, const float *x_right, const uint8_t *desc,
                                       const uint8_t *claimed_in, const float *scale_factors, int num_levels,
                                       int m, const float *reproj_x, const float *reproj_y,
                                       const float *q_x_right, const int32_t *scale_level,
                                       const uint8_t *q_desc, const uint8_t *q_valid, float margin,
                                       float lowe_ratio, int32_t *best_idx_out) {
    // match/projection.cc:37-121
    (void)num_levels;
    Grid grid(g, x, y, n);
    std::vector<uint8_t> claimed(n, 0);
    if (claimed_in) claimed.assign(claimed_in, claimed_in + n);
    unsigned num_matches = 0;
    for (int q = 0; q < m; ++q) {
        best_idx_out[q] = -1;
        if (q_valid && !q_valid[q]) continue;
        const int pred_scale_level = scale_level[q];
        const auto indices = grid.query(x, y, octave, reproj_x[q], reproj_y[q],
                                        margin * scale_factors[pred_scale_level], pred_scale_level - 1,
                                        pred_scale_level);
        if (indices.empty()) continue;
        unsigned best_hamm_dist = MAX_HAMMING_DIST, second_best_hamm_dist = MAX_HAMMING_DIST;
        int best_scale_level = -1, second_best_scale_level = -1, best_idx = -1;
        for (const auto idx : indices) {
            if (claimed[idx]) continue;
            if (x_right && 0 < x_right[idx]) {
                const auto reproj_error = std::abs(q_x_right[q] - x_right[idx]);
                if (margin * scale_factors[pred_scale_level] < reproj_error) continue;
            }
            const auto dist = orc_hamming_32(q_desc + 32 * q, desc + 32 * idx);
            if (dist < best_hamm_dist) {
                second_best_hamm_dist = best_hamm_dist;
                best_hamm_dist = dist;
                second_best_scale_level = best_scale_level;
                best_scale_level = octave[idx];
                best_idx = idx;
            } else if (dist < second_best_hamm_dist) {
                second_best_scale_level = octave[idx];
                second_best_hamm_dist = dist;
            }
        }
        if (best_hamm_dist <= HAMMING_DIST_THR_HIGH) {
            if (best_scale_level == second_best_scale_level && best_hamm_dist > lowe_ratio * second_best_hamm_dist)
                continue;
            best_idx_out[q] = best_idx;
            claimed[best_idx] = 1;  // frm.landmarks_.at(best_idx) = local_lm (local landmarks have observations)
            ++num_matches;
        }
    }
    return num_matches;
}

unsigned orc_match_current_and_last_frames(const orc_grid *g, int n, const float *x, const float *y,
                                           const int32_t *octave, const float *angle, const float *x_right,
                                           const uint8_t *desc, const uint8_t *claimed_in,
                                           const float *scale_factors, int num_levels, const orc_camera *cam,
                                           const double *pose_cw_curr, const double *pose_cw_last, int n_last,
                                           const double *pos_w, const int32_t *last_octave,
                                           const float *last_angle, const uint8_t *last_desc,
                                           const uint8_t *last_valid, float margin, int check_orientation,
                                           int32_t *matched_last_idx_out) {
    // match/projection.cc:214-358
    Grid grid(g, x, y, n);
    std::vector<uint8_t> claimed(n, 0);
    if (claimed_in) claimed.assign(claimed_in, claimed_in + n);
    for (int i = 0; i < n; ++i) matched_last_idx_out[i] = -1;
    const Pose cw(pose_cw_curr), lw(pose_cw_last);
    bool assume_forward, assume_backward;
    motion_assumption(cam, cw, lw, &assume_forward, &assume_backward);
    AngleChecker angle_checker;
    unsigned num_matches = 0;
    for (int idx_last = 0; idx_last < n_last; ++idx_last) {
        if (last_valid && !last_valid[idx_last]) continue;
        double reproj[2];
        float xr;
        if (!orc_reproject_to_image(cam, cw.R, cw.t, pos_w + 3 * idx_last, reproj, &xr)) continue;
        const int last_scale_level = last_octave[idx_last];
        const float radius = margin * scale_factors[last_scale_level];
        std::vector<unsigned> indices;
        if (assume_forward)
            indices = grid.query(x, y, octave, (float)reproj[0], (float)reproj[1], radius, last_scale_level,
                                 num_levels - 1);
        else if (assume_backward)
            indices = grid.query(x, y, octave, (float)reproj[0], (float)reproj[1], radius, 0, last_scale_level);
        else
            indices = grid.query(x, y, octave, (float)reproj[0], (float)reproj[1], radius, last_scale_level - 1,
                                 last_scale_level + 1);
        if (indices.empty()) continue;
        unsigned best_hamm_dist = MAX_HAMMING_DIST;
        int best_idx = -1;
        for (const auto curr_idx : indices) {
            if (claimed[curr_idx]) continue;
            if (x_right && x_right[curr_idx] > 0) {
                const float reproj_error = std::fabs(xr - x_right[curr_idx]);
                if (radius < reproj_error) continue;
            }
            const auto hamm_dist = orc_hamming_32(last_desc + 32 * idx_last, desc + 32 * curr_idx);
            if (hamm_dist < best_hamm_dist) {
                best_hamm_dist = hamm_dist;
                best_idx = curr_idx;
            }
        }
        if (HAMMING_DIST_THR_HIGH < best_hamm_dist) continue;
        matched_last_idx_out[best_idx] = idx_last;
        claimed[best_idx] = 1;
        ++num_matches;
        if (check_orientation) {
            const auto delta_angle = last_angle[idx_last] - angle[best_idx];
            angle_checker.append(delta_angle, best_idx);
        }
    }
    if (check_orientation) {
        for (const auto invalid_idx : angle_checker.collect(false)) {
            matched_last_idx_out[invalid_idx] = -1;
            --num_matches;
        }
    }
    return num_matches;
}

// data/landmark.cc:319-340 / data/landmark_line.cc:366-387
static unsigned predict_scale_level(float max_valid_dist, float cam_to_lm_dist, float log_scale_factor, unsigned num_levels) {
    const float ratio = max_valid_dist / cam_to_lm_dist;
    const int pred = static_cast<int>(std::ceil(std::log(ratio) / log_scale_factor));
    if (pred < 0) return 0;
    if (num_levels <= static_cast<unsigned>(pred)) return num_levels - 1;
    return static_cast<unsigned>(pred);
}

static void cam_center_of(const Pose &cw, double c[3]) {
    for (int r = 0; r < 3; ++r) c[r] = -(cw.R[0 * 3 + r] * cw.t[0] + cw.R[1 * 3 + r] * cw.t[1] + cw.R[2 * 3 + r] * cw.t[2]);
}

unsigned orc_match_frame_and_keyframe(const orc_grid *g, int n, const float *x, const float *y, const int32_t *octave,
                                      const float *angle, const uint8_t *desc, const uint8_t *claimed_in,
                                      const float *scale_factors, int num_levels, float log_scale_factor,
                                      const orc_camera *cam, const double *pose_cw_curr, int n_kf, const double *pos_w,
                                      const float *min_valid_dist, const float *max_valid_dist, const float *kf_angle,
                                      const uint8_t *kf_desc, const uint8_t *kf_valid, float margin, unsigned hamm_dist_thr,
                                      int check_orientation, int32_t *matched_kf_idx_out, float *q_reproj_x,
                                      float *q_reproj_y, int32_t *q_level, uint8_t *q_valid) {
    // match/projection.cc:529-645
    Grid grid(g, x, y, n);
    std::vector<uint8_t> claimed(n, 0);
    if (claimed_in) claimed.assign(claimed_in, claimed_in + n);
    for (int i = 0; i < n; ++i) matched_kf_idx_out[i] = -1;
    const Pose cw(pose_cw_curr);
    double cc[3];
    cam_center_of(cw, cc);
    AngleChecker angle_checker;
    unsigned num_matches = 0;
    for (int idx = 0; idx < n_kf; ++idx) {
        if (q_valid) {
            q_valid[idx] = 0;
            q_reproj_x[idx] = q_reproj_y[idx] = 0.f;
            q_level[idx] = 0;
        }
        if (kf_valid && !kf_valid[idx]) continue;  // !lm || will_be_erased || already_matched
        double reproj[2];
        float xr;
        if (!orc_reproject_to_image(cam, cw.R, cw.t, pos_w + 3 * idx, reproj, &xr)) continue;
        const double v[3] = {pos_w[3 * idx] - cc[0], pos_w[3 * idx + 1] - cc[1], pos_w[3 * idx + 2] - cc[2]};
        const double dist = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        const float max_d = (float)(1.3 * max_valid_dist[idx]), min_d = (float)(0.7 * min_valid_dist[idx]);  // landmark.cc:297-307
        if (dist < min_d || max_d < dist) continue;
        const unsigned lvl = predict_scale_level(max_valid_dist[idx], (float)dist, log_scale_factor, (unsigned)num_levels);
        if (q_valid) {
            q_valid[idx] = 1;
            q_reproj_x[idx] = (float)reproj[0];
            q_reproj_y[idx] = (float)reproj[1];
            q_level[idx] = (int)lvl;
        }
        const auto indices = grid.query(x, y, octave, (float)reproj[0], (float)reproj[1], margin * scale_factors[lvl],
                                        (int)lvl - 1, (int)lvl + 1);
        if (indices.empty()) continue;
        unsigned best_hamm_dist = MAX_HAMMING_DIST;
        int best_idx = -1;
        for (const auto curr_idx : indices) {
            if (claimed[curr_idx]) continue;
            const auto hamm_dist = orc_hamming_32(kf_desc + 32 * idx, desc + 32 * curr_idx);
            if (hamm_dist < best_hamm_dist) {
                best_hamm_dist = hamm_dist;
                best_idx = curr_idx;
            }
        }
        if (hamm_dist_thr < best_hamm_dist) continue;
        matched_kf_idx_out[best_idx] = idx;
        claimed[best_idx] = 1;
        ++num_matches;
        if (check_orientation) angle_checker.append(kf_angle[idx] - angle[best_idx], best_idx);
    }
    if (check_orientation) {
        for (const auto invalid_idx : angle_checker.collect(false)) {
            matched_kf_idx_out[invalid_idx] = -1;
            --num_matches;
        }
    }
    return num_matches;
}

unsigned orc_match_frame_and_keyframe_line(int n, const float *sx, const float *sy, const float *ex, const float *ey,
                                           const int32_t *octave, const uint8_t *desc, const uint8_t *claimed_in,
                                           const float *scale_factors_lsd, int num_levels_lsd, float log_scale_factor_lsd,
                                           const orc_camera *cam, const double *pose_cw_curr, int n_kf,
                                           const double *pos_w /*n_kf x 6*/, const float *min_valid_dist,
                                           const float *max_valid_dist, const uint8_t *kf_desc, const uint8_t *kf_valid,
                                           float margin, unsigned hamm_dist_thr, int32_t *matched_kf_idx_out,
                                           float *q_sp_x, float *q_sp_y, float *q_ep_x, float *q_ep_y, int32_t *q_level,
                                           uint8_t *q_valid) {
    // match/projection.cc:648-779
    std::vector<uint8_t> claimed(n, 0);
    if (claimed_in) claimed.assign(claimed_in, claimed_in + n);
    for (int i = 0; i < n; ++i) matched_kf_idx_out[i] = -1;
    const Pose cw(pose_cw_curr);
    double cc[3];
    cam_center_of(cw, cc);
    unsigned num_matches = 0;
    for (int idx = 0; idx < n_kf; ++idx) {
        if (q_valid) {
            q_valid[idx] = 0;
            q_sp_x[idx] = q_sp_y[idx] = q_ep_x[idx] = q_ep_y[idx] = 0.f;
            q_level[idx] = 0;
        }
        if (kf_valid && !kf_valid[idx]) continue;
        const double *p = pos_w + 6 * idx;
        double rsp[2], rep[2], rmp[2];
        float xr;
        const bool in_sp = orc_reproject_to_image(cam, cw.R, cw.t, p, rsp, &xr) != 0;
        const bool in_ep = orc_reproject_to_image(cam, cw.R, cw.t, p + 3, rep, &xr) != 0;
        if (!in_sp && !in_ep) continue;
        const double mp[3] = {0.5 * (p[0] + p[3]), 0.5 * (p[1] + p[4]), 0.5 * (p[2] + p[5])};
        if (!in_sp || !in_ep) {
            if (!orc_reproject_to_image(cam, cw.R, cw.t, mp, rmp, &xr)) continue;
        }
        const double v[3] = {mp[0] - cc[0], mp[1] - cc[1], mp[2] - cc[2]};
        const double dist = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        const float max_d = (float)(1.3 * max_valid_dist[idx]), min_d = (float)(0.7 * min_valid_dist[idx]);
        if (dist < min_d || max_d < dist) continue;
        const unsigned lvl = predict_scale_level(max_valid_dist[idx], (float)dist, log_scale_factor_lsd, (unsigned)num_levels_lsd);
        if (q_valid) {
            q_valid[idx] = 1;
            q_sp_x[idx] = (float)rsp[0];
            q_sp_y[idx] = (float)rsp[1];
            q_ep_x[idx] = (float)rep[0];
            q_ep_y[idx] = (float)rep[1];
            q_level[idx] = (int)lvl;
        }
        const auto indices = keylines_in_cell(n, sx, sy, ex, ey, octave, (float)rsp[0], (float)rsp[1], (float)rep[0],
                                              (float)rep[1], margin * scale_factors_lsd[lvl], (int)lvl - 1, (int)lvl + 1);
        if (indices.empty()) continue;
        unsigned best_hamm_dist = MAX_HAMMING_DIST;
        int best_idx = -1;
        for (const auto curr_idx : indices) {
            if (claimed[curr_idx]) continue;
            const auto hamm_dist = orc_hamming_32(kf_desc + 32 * idx, desc + 32 * curr_idx);
            if (hamm_dist < best_hamm_dist) {
                best_hamm_dist = hamm_dist;
                best_idx = curr_idx;
            }
        }
        if (hamm_dist_thr < best_hamm_dist) continue;
        matched_kf_idx_out[best_idx] = idx;
        claimed[best_idx] = 1;
        ++num_matches;
    }
    return num_matches;
}

// match/robust.cc:387-406.  libm != 0: std::acos as the reference writes it; 0: acos(c) = atan2(sqrt((1-c)(1+c)), c) with
// the deterministic kernel of detmath.h (what the CUDA path evaluates)
static bool check_epipolar_constraint(const double *b1, const double *b2, const double *E, float sf1, int libm) {
    const double e0 = E[0] * b2[0] + E[1] * b2[1] + E[2] * b2[2];
    const double e1 = E[3] * b2[0] + E[4] * b2[1] + E[5] * b2[2];
    const double e2 = E[6] * b2[0] + E[7] * b2[1] + E[8] * b2[2];
    const double cos_residual = (e0 * b1[0] + e1 * b1[1] + e2 * b1[2]) / std::sqrt(e0 * e0 + e1 * e1 + e2 * e2);
    const double ac = libm ? std::acos(cos_residual)
                           : det_atan2(std::sqrt((1.0 - cos_residual) * (1.0 + cos_residual)), cos_residual);
    const double residual_rad = M_PI / 2.0 - std::abs(ac);
    constexpr double residual_rad_thr = 0.2 * M_PI / 180.0;
    return residual_rad < residual_rad_thr * sf1;
}

unsigned orc_match_for_triangulation(int n1, const uint8_t *desc1, const float *angle1, const int32_t *octave1,
                                     const double *bearing1, const uint8_t *has_lm1, const float *x_right1, int n2,
                                     const uint8_t *desc2, const float *angle2, const double *bearing2,
                                     const uint8_t *has_lm2, const float *x_right2, int nodes1, const uint32_t *ids1,
                                     const int32_t *off1, const uint32_t *idx1, int nodes2, const uint32_t *ids2,
                                     const int32_t *off2, const uint32_t *idx2, const double *E_12, const double *epipole,
                                     const float *scale_factors_1, int check_orientation, int libm,
                                     int32_t *matched_idx2_in_1_out) {
    // match/robust.cc:43-216
    unsigned num_matches = 0;
    AngleChecker angle_checker;
    std::vector<bool> already2(n2, false);
    for (int i = 0; i < n1; ++i) matched_idx2_in_1_out[i] = -1;
    int a = 0, b = 0;
    while (a < nodes1 && b < nodes2) {
        if (ids1[a] == ids2[b]) {
            for (int k1 = off1[a]; k1 < off1[a + 1]; ++k1) {
                const unsigned i1 = idx1[k1];
                if (has_lm1[i1]) continue;
                const bool st1 = x_right1 ? 0 <= x_right1[i1] : false;
                unsigned best_hamm_dist = HAMMING_DIST_THR_LOW;
                int best_idx_2 = -1;
                for (int k2 = off2[b]; k2 < off2[b + 1]; ++k2) {
                    const unsigned i2 = idx2[k2];
                    if (has_lm2[i2]) continue;
                    if (already2[i2]) continue;
                    const bool st2 = x_right2 ? 0 <= x_right2[i2] : false;
                    const unsigned hamm_dist = orc_hamming_32(desc1 + 32 * i1, desc2 + 32 * i2);
                    if (HAMMING_DIST_THR_LOW < hamm_dist || best_hamm_dist < hamm_dist) continue;
                    if (!st1 && !st2) {
                        const double *b2 = bearing2 + 3 * i2;
                        const double cos_dist = epipole[0] * b2[0] + epipole[1] * b2[1] + epipole[2] * b2[2];
                        constexpr double cos_dist_thr = 0.99862953475;
                        if (cos_dist_thr < cos_dist) continue;
                    }
                    if (check_epipolar_constraint(bearing1 + 3 * i1, bearing2 + 3 * i2, E_12,
                                                  scale_factors_1[octave1[i1]], libm)) {
                        best_idx_2 = (int)i2;
                        best_hamm_dist = hamm_dist;
                    }
                }
                if (best_idx_2 < 0) continue;
                already2[best_idx_2] = true;
                matched_idx2_in_1_out[i1] = best_idx_2;
                ++num_matches;
                if (check_orientation) angle_checker.append(angle1[i1] - angle2[best_idx_2], (int)i1);
            }
            ++a;
            ++b;
        } else if (ids1[a] < ids2[b]) {
            while (a < nodes1 && ids1[a] < ids2[b]) ++a;  // lower_bound(itr_2->first)
        } else {
            while (b < nodes2 && ids2[b] < ids1[a]) ++b;
        }
    }
    if (check_orientation) {
        for (const auto invalid_idx : angle_checker.collect(false)) {
            matched_idx2_in_1_out[invalid_idx] = -1;
            --num_matches;
        }
    }
    return num_matches;
}

void orc_landmark_compute_descriptor_batch(const uint8_t *descs, const int32_t *offsets, int num_landmarks,
                                           int32_t *best_idx_out) {
    // data/landmark.cc:181-247
    for (int l = 0; l < num_landmarks; ++l) {
        const int beg = offsets[l], num_descs = offsets[l + 1] - beg;
        if (num_descs <= 0) {
            best_idx_out[l] = -1;
            continue;
        }
        std::vector<std::vector<unsigned>> hamm_dists(num_descs, std::vector<unsigned>(num_descs));
        for (int i = 0; i < num_descs; ++i) {
            hamm_dists[i][i] = 0;
            for (int j = i + 1; j < num_descs; ++j) {
                const auto dist = orc_hamming_32(descs + 32 * (size_t)(beg + i), descs + 32 * (size_t)(beg + j));
                hamm_dists[i][j] = dist;
                hamm_dists[j][i] = dist;
            }
        }
        unsigned best_median_dist = MAX_HAMMING_DIST, best_idx = 0;
        for (int idx = 0; idx < num_descs; ++idx) {
            std::vector<unsigned> partial(hamm_dists[idx].begin(), hamm_dists[idx].begin() + num_descs);
            std::sort(partial.begin(), partial.end());
            const auto median_dist = partial.at(static_cast<unsigned>(0.5 * (num_descs - 1)));
            if (median_dist < best_median_dist) {
                best_median_dist = median_dist;
                best_idx = idx;
            }
        }
        best_idx_out[l] = (int)best_idx;
    }
}

unsigned orc_match_frame_and_landmarks_line(int n, const float *sx, const float *sy, const float *ex,
                                            const float *ey, const int32_t *octave,
                                            const int32_t *ratio_level, const uint8_t *desc,
                                            const uint8_t *claimed_in, const float *scale_factors_lsd,
                                            int num_levels_lsd, int m, const float *sp_x, const float *sp_y,
                                            const float *ep_x, const float *ep_y, const int32_t *scale_level,
                                            const uint8_t *q_desc, const uint8_t *q_valid, float margin,
                                            float lowe_ratio, int32_t *best_idx_out) {
    // match/projection.cc:124-212
    (void)num_levels_lsd;
    std::vector<uint8_t> claimed(n, 0);
    if (claimed_in) claimed.assign(claimed_in, claimed_in + n);
    unsigned num_matches = 0;
    for (int q = 0; q < m; ++q) {
        best_idx_out[q] = -1;
        if (q_valid && !q_valid[q]) continue;
        const int pred_scale_level = scale_level[q];
        const auto indices =
            keylines_in_cell(n, sx, sy, ex, ey, octave, sp_x[q], sp_y[q], ep_x[q], ep_y[q],
                             margin * scale_factors_lsd[pred_scale_level], pred_scale_level - 1, pred_scale_level);
        if (indices.empty()) continue;
        unsigned best_hamm_dist = MAX_HAMMING_DIST, second_best_hamm_dist = MAX_HAMMING_DIST;
        int best_scale_level = -1, second_best_scale_level = -1, best_idx = -1;
        for (const auto idx : indices) {
            if (claimed[idx]) continue;
            const auto dist = orc_hamming_32(q_desc + 32 * q, desc + 32 * idx);
            if (dist < best_hamm_dist) {
                second_best_hamm_dist = best_hamm_dist;
                best_hamm_dist = dist;
                second_best_scale_level = best_scale_level;
                best_scale_level = ratio_level[idx];
                best_idx = idx;
            } else if (dist < second_best_hamm_dist) {
                second_best_scale_level = ratio_level[idx];
                second_best_hamm_dist = dist;
            }
        }
        if (best_hamm_dist <= HAMMING_DIST_THR_HIGH) {
            if (best_scale_level == second_best_scale_level && best_hamm_dist > lowe_ratio * second_best_hamm_dist)
                continue;
            best_idx_out[q] = best_idx;
            claimed[best_idx] = 1;
            ++num_matches;
        }
    }
    return num_matches;
}

unsigned orc_match_current_and_last_frames_line(int n, const float *sx, const float *sy, const float *ex,
                                                const float *ey, const int32_t *octave,
                                                const float *x_right_sp, const float *x_right_ep,
                                                const uint8_t *desc, const uint8_t *claimed_in,
                                                const float *scale_factors_lsd, int num_levels_lsd,
                                                const orc_camera *cam, const double *pose_cw_curr,
                                                const double *pose_cw_last, int n_last, const double *pos_w,
                                                const int32_t *last_octave, const uint8_t *last_desc,
                                                const uint8_t *last_valid, float margin,
                                                int32_t *matched_last_idx_out) {
    // match/projection.cc:361-527
    std::vector<uint8_t> claimed(n, 0);
    if (claimed_in) claimed.assign(claimed_in, claimed_in + n);
    for (int i = 0; i < n; ++i) matched_last_idx_out[i] = -1;
    const Pose cw(pose_cw_curr), lw(pose_cw_last);
    bool assume_forward, assume_backward;
    motion_assumption(cam, cw, lw, &assume_forward, &assume_backward);
    unsigned num_matches = 0;
    for (int idx_last = 0; idx_last < n_last; ++idx_last) {
        if (last_valid && !last_valid[idx_last]) continue;
        const double *pw = pos_w + 6 * idx_last;
        double reproj_sp[2], reproj_ep[2];
        float xr_sp, xr_ep;
        const bool in_sp = orc_reproject_to_image(cam, cw.R, cw.t, pw, reproj_sp, &xr_sp);
        const bool in_ep = orc_reproject_to_image(cam, cw.R, cw.t, pw + 3, reproj_ep, &xr_ep);
        if (!in_sp && !in_ep) continue;
        if (!in_sp || !in_ep) {
            const double mp[3] = {0.5 * (pw[0] + pw[3]), 0.5 * (pw[1] + pw[4]), 0.5 * (pw[2] + pw[5])};
            double reproj_mp[2];
            float xr_mp;
            if (!orc_reproject_to_image(cam, cw.R, cw.t, mp, reproj_mp, &xr_mp)) continue;
        }
        // NB: when an endpoint is behind the camera (z<=0) the reference leaves reproj_* uninitialised
        // (camera/perspective.cc:196-199); the oracle and the kernels define it as 0 in that case.
        const int last_scale_level = last_octave[idx_last];
        const float radius = margin * scale_factors_lsd[last_scale_level];
        int min_level, max_level;
        if (assume_forward) {
            min_level = last_scale_level;
            max_level = num_levels_lsd;
        } else if (assume_backward) {
            min_level = 0;
            max_level = last_scale_level + 1;
        } else {
            min_level = last_scale_level - 1;
            max_level = last_scale_level + 1;
        }
        const auto indices = keylines_in_cell(n, sx, sy, ex, ey, octave, (float)reproj_sp[0], (float)reproj_sp[1],
                                              (float)reproj_ep[0], (float)reproj_ep[1], radius, min_level, max_level);
        if (indices.empty()) continue;
        unsigned best_hamm_dist = MAX_HAMMING_DIST;
        int best_idx = -1;
        for (const auto curr_idx : indices) {
            if (claimed[curr_idx]) continue;
            if (cam->setup_type == 2 && x_right_sp && x_right_ep) {
                if (x_right_sp[curr_idx] > 0 && x_right_ep[curr_idx] > 0) {
                    const float e_sp = std::fabs(xr_sp - x_right_sp[curr_idx]);
                    const float e_ep = std::fabs(xr_ep - x_right_ep[curr_idx]);
                    if (radius < e_sp || radius < e_ep) continue;
                }
            }
            const auto hamm_dist = orc_hamming_32(last_desc + 32 * idx_last, desc + 32 * curr_idx);
            if (hamm_dist < best_hamm_dist) {
                best_hamm_dist = hamm_dist;
                best_idx = curr_idx;
            }
        }
        if (HAMMING_DIST_THR_HIGH < best_hamm_dist) continue;
        matched_last_idx_out[best_idx] = idx_last;
        claimed[best_idx] = 1;
        ++num_matches;
    }
    return num_matches;
}

// =========================================================================================
// match/fuse.cc -- per-landmark search of the fuse matchers (the effects stay with the caller)
// =========================================================================================
unsigned orc_predict_scale_level(float max_valid_dist, float cam_to_lm_dist, float log_scale_factor, unsigned num_levels) {
    return predict_scale_level(max_valid_dist, cam_to_lm_dist, log_scale_factor, num_levels);
}

void orc_fuse_search_points(const orc_grid *g, const orc_camera *cam, int n, const float *x, const float *y,
                            const int32_t *octave, const float *x_right, const uint8_t *desc, const double *rot_cw,
                            const double *trans_cw, const double *cam_center, const float *scale_factors,
                            const float *inv_level_sigma_sq, int num_levels, float log_scale_factor, int m,
                            const double *pos_w, const double *obs_mean_normal, const float *min_valid_dist,
                            const float *max_valid_dist, const float *max_valid_dist_raw, const uint8_t *lm_desc,
                            const uint8_t *lm_valid, const uint8_t *lm_skip, float margin, int mode,
                            int32_t *best_idx_out, uint16_t *best_dist_out, int32_t *level_out) {
    Grid grid(g, x, y, n);
    for (int i = 0; i < m; ++i) {
        best_idx_out[i] = -1;
        if (best_dist_out) best_dist_out[i] = 0xFFFF;
        if (level_out) level_out[i] = -1;
        if (lm_valid && !lm_valid[i]) continue;  // fuse.cc:163-170 / 58-61: !lm || will_be_erased
        if (lm_skip && lm_skip[i]) continue;     // :171-174 is_observed_in_keyframe / :63-66 valid_lms_in_keyfrm.count
        const double *pw = pos_w + 3 * i;
        // :180-188 / 72-80
        double reproj[2];
        float xr;
        if (!orc_reproject_to_image(cam, rot_cw, trans_cw, pw, reproj, &xr)) continue;
        // :190-199 / 82-91
        const double v[3] = {pw[0] - cam_center[0], pw[1] - cam_center[1], pw[2] - cam_center[2]};
        const double cam_to_lm_dist = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        const float max_d = max_valid_dist[i], min_d = min_valid_dist[i];
        if (cam_to_lm_dist < min_d || max_d < cam_to_lm_dist) continue;
        // :201-208 / 93-100: angle to the mean observation direction below 60 deg
        const double *nm = obs_mean_normal + 3 * i;
        if (v[0] * nm[0] + v[1] * nm[1] + v[2] * nm[2] < 0.5 * cam_to_lm_dist) continue;
        // :210-218 / 102-110
        const unsigned pred = predict_scale_level(max_valid_dist_raw[i], (float)cam_to_lm_dist, log_scale_factor, (unsigned)num_levels);
        if (level_out) level_out[i] = (int)pred;
        const auto indices = grid.query(x, y, octave, (float)reproj[0], (float)reproj[1], margin * scale_factors[pred], -1, -1);
        if (indices.empty()) continue;
        unsigned best_dist = MAX_HAMMING_DIST;
        int best_idx = -1;
        for (const auto idx : indices) {
            if (mode == 0) {
                // detect_duplication :113-121 -- int arithmetic
                const int scale_level = octave[idx];
                const int pred_i = (int)pred;
                if (scale_level < pred_i - 1 || pred_i < scale_level) continue;
            } else {
                // replace_duplication :228-236 -- UNSIGNED arithmetic: pred == 0 makes pred - 1 wrap and rejects everything
                const unsigned scale_level = (unsigned)octave[idx];
                if (scale_level < pred - 1u || pred < scale_level) continue;
                const float kxr = x_right ? x_right[idx] : -1.0f;
                if (kxr >= 0) {
                    // :238-251
                    const double e_x = reproj[0] - x[idx];
                    const double e_y = reproj[1] - y[idx];
                    const float e_x_right = xr - kxr;
                    const double reproj_error_sq = e_x * e_x + e_y * e_y + e_x_right * e_x_right;
                    constexpr float chi_sq_3D = 7.81473;
                    if (chi_sq_3D < reproj_error_sq * inv_level_sigma_sq[scale_level]) continue;
                } else {
                    // :252-265
                    const double e_x = reproj[0] - x[idx];
                    const double e_y = reproj[1] - y[idx];
                    const double reproj_error_sq = e_x * e_x + e_y * e_y;
                    constexpr float chi_sq_2D = 5.99146;
                    if (chi_sq_2D < reproj_error_sq * inv_level_sigma_sq[scale_level]) continue;
                }
            }
            const auto hamm_dist = orc_hamming_32(lm_desc + 32 * i, desc + 32 * idx);
            if (hamm_dist < best_dist) {
                best_dist = hamm_dist;
                best_idx = idx;
            }
        }
        if (HAMMING_DIST_THR_LOW < best_dist) continue;  // :279-282 / 135-138
        best_idx_out[i] = best_idx;
        if (best_dist_out) best_dist_out[i] = (uint16_t)best_dist;
    }
}

void orc_fuse_search_lines(const orc_camera *cam, int n, const float *sx, const float *sy, const float *ex, const float *ey,
                           const int32_t *octave, const uint8_t *desc, const double *rot_cw, const double *trans_cw,
                           const double *cam_center, const float *scale_factors_lsd, const float *inv_level_sigma_sq_lsd,
                           int num_levels_lsd, float log_scale_factor_lsd, int m, const double *pos_w,
                           const float *min_valid_dist, const float *max_valid_dist, const float *max_valid_dist_raw,
                           const uint8_t *lm_desc, const uint8_t *lm_valid, const uint8_t *lm_skip, float margin,
                           int32_t *best_idx_out, uint16_t *best_dist_out, int32_t *level_out) {
    for (int i = 0; i < m; ++i) {
        best_idx_out[i] = -1;
        if (best_dist_out) best_dist_out[i] = 0xFFFF;
        if (level_out) level_out[i] = -1;
        if (lm_valid && !lm_valid[i]) continue;  // fuse.cc:314-321
        if (lm_skip && lm_skip[i]) continue;     // :322-325
        const double *sp = pos_w + 6 * i, *ep = sp + 3;
        // :332-339
        // a point behind the camera leaves its reprojection unset in the reference (camera/perspective.cc:197-200 returns
        // before writing); the oracle and the CUDA path define it as (0, 0)
        double rsp[2] = {0.0, 0.0}, rep[2] = {0.0, 0.0}, rmp[2];
        float xr;
        const bool in_sp = orc_reproject_to_image(cam, rot_cw, trans_cw, sp, rsp, &xr) != 0;
        const bool in_ep = orc_reproject_to_image(cam, rot_cw, trans_cw, ep, rep, &xr) != 0;
        if (!in_sp && !in_ep) continue;
        // :347-366 partial occlusion: the mid point must be visible
        const double mp[3] = {0.5 * (sp[0] + ep[0]), 0.5 * (sp[1] + ep[1]), 0.5 * (sp[2] + ep[2])};
        if (!in_sp || !in_ep) {
            if (!orc_reproject_to_image(cam, rot_cw, trans_cw, mp, rmp, &xr)) continue;
        }
        // :368-383
        const double vs[3] = {sp[0] - cam_center[0], sp[1] - cam_center[1], sp[2] - cam_center[2]};
        const double ve[3] = {ep[0] - cam_center[0], ep[1] - cam_center[1], ep[2] - cam_center[2]};
        const double dist_sp = std::sqrt(vs[0] * vs[0] + vs[1] * vs[1] + vs[2] * vs[2]);
        const double dist_ep = std::sqrt(ve[0] * ve[0] + ve[1] * ve[1] + ve[2] * ve[2]);
        const float max_d = max_valid_dist[i], min_d = min_valid_dist[i];
        if (dist_sp < min_d || max_d < dist_sp || dist_ep < min_d || max_d < dist_ep) continue;
        // :385-392
        const double vm[3] = {mp[0] - cam_center[0], mp[1] - cam_center[1], mp[2] - cam_center[2]};
        const double dist_mp = std::sqrt(vm[0] * vm[0] + vm[1] * vm[1] + vm[2] * vm[2]);
        const unsigned pred = predict_scale_level(max_valid_dist_raw[i], (float)dist_mp, log_scale_factor_lsd, (unsigned)num_levels_lsd);
        if (level_out) level_out[i] = (int)pred;
        const auto indices = keylines_in_cell(n, sx, sy, ex, ey, octave, (float)rsp[0], (float)rsp[1], (float)rep[0],
                                              (float)rep[1], margin * scale_factors_lsd[pred], -1, -1);
        if (indices.empty()) continue;
        // :405-445
        const double l0 = rsp[1] * 1.0 - 1.0 * rep[1];
        const double l1 = 1.0 * rep[0] - rsp[0] * 1.0;
        const double l2 = rsp[0] * rep[1] - rsp[1] * rep[0];
        unsigned best_dist = MAX_HAMMING_DIST;
        int best_idx = -1;
        for (const auto idx : indices) {
            const unsigned scale_level = (unsigned)octave[idx];
            const double den = std::sqrt(l0 * l0 + l1 * l1);
            const double e_sp = (sx[idx] * l0 + sy[idx] * l1 + l2) / den;
            const double e_ep = (ex[idx] * l0 + ey[idx] * l1 + l2) / den;
            constexpr float chi_sq_2D = 5.99146;
            if (chi_sq_2D < (e_sp * e_sp + e_ep * e_ep) * inv_level_sigma_sq_lsd[scale_level]) continue;
            const auto hamm_dist = orc_hamming_32(lm_desc + 32 * i, desc + 32 * idx);
            if (hamm_dist < best_dist) {
                best_dist = hamm_dist;
                best_idx = idx;
            }
        }
        if (HAMMING_DIST_THR_LOW < best_dist) continue;  // :447-450
        best_idx_out[i] = best_idx;
        if (best_dist_out) best_dist_out[i] = (uint16_t)best_dist;
    }
}

// =========================================================================================
// match/bow_tree.cc:41-165 / :167-305
// =========================================================================================
unsigned orc_bow_tree_match(int n1, const uint8_t *desc1, const float *angle1, const uint8_t *valid1, int n2,
                            const uint8_t *desc2, const float *angle2, const uint8_t *valid2, int nodes1,
                            const uint32_t *ids1, const int32_t *off1, const uint32_t *idx1, int nodes2,
                            const uint32_t *ids2, const int32_t *off2, const uint32_t *idx2, float lowe_ratio,
                            int check_orientation, int32_t *matched_2_of_1, int32_t *matched_1_of_2) {
    unsigned num_matches = 0;
    AngleChecker angle_checker;
    for (int i = 0; i < n1; ++i) matched_2_of_1[i] = -1;
    std::vector<int32_t> m1of2(n2, -1);  // matched_lms_in_frm / is_already_matched_in_keyfrm_2
    int a = 0, b = 0;
    while (a < nodes1 && b < nodes2) {
        if (ids1[a] == ids2[b]) {
            for (int ka = off1[a]; ka < off1[a + 1]; ++ka) {
                const int i1 = (int)idx1[ka];
                if (valid1 && !valid1[i1]) continue;  // :70-78 / :219-227
                unsigned best_hamm_dist = MAX_HAMMING_DIST, second_best_hamm_dist = MAX_HAMMING_DIST;
                int best_idx_2 = -1;
                for (int kb = off2[b]; kb < off2[b + 1]; ++kb) {
                    const int i2 = (int)idx2[kb];
                    if (valid2 && !valid2[i2]) continue;  // :238-247 (match_keyframes only)
                    if (m1of2[i2] >= 0) continue;         // :89-92 / :249-252
                    const auto hamm_dist = orc_hamming_32(desc1 + 32 * (size_t)i1, desc2 + 32 * (size_t)i2);
                    if (hamm_dist < best_hamm_dist) {
                        second_best_hamm_dist = best_hamm_dist;
                        best_hamm_dist = hamm_dist;
                        best_idx_2 = i2;
                    } else if (hamm_dist < second_best_hamm_dist) {
                        second_best_hamm_dist = hamm_dist;
                    }
                }
                if (HAMMING_DIST_THR_LOW < best_hamm_dist) continue;                                  // :110-113
                if (lowe_ratio * second_best_hamm_dist < static_cast<float>(best_hamm_dist)) continue;  // :115-119
                matched_2_of_1[i1] = best_idx_2;
                m1of2[best_idx_2] = i1;
                if (check_orientation) angle_checker.append(angle1[i1] - angle2[best_idx_2], i1);  // :123-127
                ++num_matches;
            }
            ++a;
            ++b;
        } else if (ids1[a] < ids2[b]) {
            ++a;  // lower_bound on an ascending map
        } else {
            ++b;
        }
    }
    if (check_orientation) {
        for (const auto invalid_i1 : angle_checker.collect(false)) {  // :152-160
            m1of2[matched_2_of_1[invalid_i1]] = -1;
            matched_2_of_1[invalid_i1] = -1;
            --num_matches;
        }
    }
    if (matched_1_of_2)
        for (int j = 0; j < n2; ++j) matched_1_of_2[j] = m1of2[j];
    return num_matches;
}

unsigned orc_brute_force_match(const uint8_t *frm_desc, const float *frm_angle, int n_frm,
                               const uint8_t *kf_desc, const float *kf_angle, const uint8_t *kf_valid, int n_kf,
                               float lowe_ratio, int check_orientation, int32_t *matched) {
    // match/robust.cc:257-385
    unsigned num_matches = 0;
    AngleChecker angle_checker;
    for (int i = 0; i < n_frm; ++i) matched[i] = -1;
    std::vector<uint8_t> already(n_frm, 0);
    for (int idx_2 = 0; idx_2 < n_kf; ++idx_2) {
        if (kf_valid && !kf_valid[idx_2]) continue;
        unsigned best_hamm_dist = MAX_HAMMING_DIST, second_best_hamm_dist = MAX_HAMMING_DIST;
        int best_idx_1 = -1;
        for (int idx_1 = 0; idx_1 < n_frm; ++idx_1) {
            if (already[idx_1]) continue;
            const auto hamm_dist = orc_hamming_32(kf_desc + 32 * idx_2, frm_desc + 32 * idx_1);
            if (hamm_dist < best_hamm_dist) {
                second_best_hamm_dist = best_hamm_dist;
                best_hamm_dist = hamm_dist;
                best_idx_1 = idx_1;
            } else if (hamm_dist < second_best_hamm_dist) {
                second_best_hamm_dist = hamm_dist;
            }
        }
        if (HAMMING_DIST_THR_LOW < best_hamm_dist) continue;
        if (best_idx_1 < 0) continue;
        if (lowe_ratio * second_best_hamm_dist < static_cast<float>(best_hamm_dist)) continue;
        matched[best_idx_1] = idx_2;
        already[best_idx_1] = 1;
        if (check_orientation) {
            const auto delta_angle = frm_angle[best_idx_1] - kf_angle[idx_2];
            angle_checker.append(delta_angle, best_idx_1);
        }
        ++num_matches;
    }
    if (check_orientation) {
        for (const auto invalid_idx_1 : angle_checker.collect(false)) {
            matched[invalid_idx_1] = -1;
            --num_matches;
        }
    }
    return num_matches;
}

}  // extern "C"
