// plpslam_b200_adapter.hpp -- reference-side adapters: marshal the reference's own types (cv::Mat,
// cv::KeyPoint, data::frame, data::landmark, g2o-free PODs) into the C ABI of include/plpslam_b200.h.
//
// This header is compiled INSIDE the reference tree (it needs OpenCV / Eigen / the PLPSLAM headers, none of which
// exist in the authoring environment) and is what the patched bodies of
//   src/PLPSLAM/feature/orb_extractor.cc        (orb_extractor::extract)
//   src/PLPSLAM/match/projection.cc             (match_frame_and_landmarks, match_current_and_last_frames, *_line)
//   src/PLPSLAM/match/robust.cc                 (brute_force_match)
//   src/PLPSLAM/optimize/pose_optimizer*.cc     (optimize)
// call.  Public signatures, PLPSLAM::system, the YAML configs and the map database stay unchanged.
// See INTEGRATION.md for the patch of each call site.
#pragma once
#ifdef PLPSLAM_B200_WITH_REFERENCE_TYPES

#include <stdexcept>
#include <vector>

#include <opencv2/core.hpp>

#include "PLPSLAM/camera/perspective.h"
#include "PLPSLAM/data/frame.h"
#include "PLPSLAM/data/landmark.h"
#include "PLPSLAM/data/landmark_line.h"
#include "plpslam_b200.h"

namespace plpslam_b200 {

inline void check(plp_status s) {
    // the reference's operators do not return errors; a failing GPU call is a hard error, never a CPU fallback
    if (s != PLP_OK) throw std::runtime_error(std::string("plpslam_b200: ") + plp_last_error());
}

// one context per calling thread (tracking thread, mapping thread), like the reference's per-thread objects
inline plp_ctx *thread_ctx(int device = 0) {
    thread_local plp_ctx *ctx = nullptr;
    if (!ctx) check(plp_ctx_create(device, &ctx));
    return ctx;
}

inline plp_grid grid_of(const PLPSLAM::camera::base *c) {
    return plp_grid{c->img_bounds_.min_x_, c->img_bounds_.min_y_, c->inv_cell_width_, c->inv_cell_height_,
                    (int32_t)c->num_grid_cols_, (int32_t)c->num_grid_rows_};
}

inline plp_camera camera_of(const PLPSLAM::camera::base *b) {
    const auto *c = static_cast<const PLPSLAM::camera::perspective *>(b);
    return plp_camera{c->fx_, c->fy_, c->cx_, c->cy_, b->focal_x_baseline_, b->true_baseline_,
                      b->img_bounds_.min_x_, b->img_bounds_.max_x_, b->img_bounds_.min_y_, b->img_bounds_.max_y_,
                      (int32_t)b->setup_type_};
}

// ---- feature::orb_extractor::extract (feature/orb_extractor.cc:73-160) -------------------------------------
struct orb_backend {
    plp_orb *h = nullptr;
    int rows = 0, cols = 0;
    std::vector<plp_keypoint> kp;
    void ensure(const PLPSLAM::feature::orb_params &p, int r, int c) {
        if (h && r == rows && c == cols) return;
        if (h) plp_orb_destroy(h);
        plp_orb_params q{p.max_num_keypts_, p.scale_factor_, p.num_levels_, p.ini_fast_thr_, p.min_fast_thr};
        check(plp_orb_create(thread_ctx(), &q, r, c, 1, &h));
        rows = r;
        cols = c;
        kp.resize(plp_orb_capacity(h));
    }
    void extract(const cv::Mat &image, const cv::Mat &mask, std::vector<cv::KeyPoint> &keypts,
                 const cv::_OutputArray &out_descriptors) {
        cv::Mat desc(plp_orb_capacity(h), 32, CV_8U);
        int n = 0;
        check(plp_orb_extract(h, image.data, image.rows, image.cols, image.step, mask.empty() ? nullptr : mask.data,
                              mask.empty() ? 0 : mask.step, kp.data(), desc.data, &n));
        static_assert(sizeof(plp_keypoint) == sizeof(cv::KeyPoint), "plp_keypoint mirrors cv::KeyPoint");
        keypts.assign(reinterpret_cast<cv::KeyPoint *>(kp.data()), reinterpret_cast<cv::KeyPoint *>(kp.data()) + n);
        if (n == 0)
            out_descriptors.release();
        else
            desc.rowRange(0, n).copyTo(out_descriptors);
    }
};

// ---- match::projection::match_frame_and_landmarks (match/projection.cc:37-121) -------------------------------
inline unsigned match_frame_and_landmarks(PLPSLAM::data::frame &frm,
                                          const std::vector<PLPSLAM::data::landmark *> &local_landmarks, float margin,
                                          float lowe_ratio) {
    const int n = frm.num_keypts_, m = (int)local_landmarks.size();
    std::vector<float> x(n), y(n), xr(n), qx(m), qy(m), qxr(m);
    std::vector<int32_t> oct(n), lvl(m), best(m);
    std::vector<uint8_t> claimed(n), valid(m), qdesc((size_t)m * 32);
    for (int i = 0; i < n; ++i) {
        x[i] = frm.undist_keypts_[i].pt.x;
        y[i] = frm.undist_keypts_[i].pt.y;
        oct[i] = frm.undist_keypts_[i].octave;
        xr[i] = frm.stereo_x_right_[i];
        claimed[i] = frm.landmarks_[i] && frm.landmarks_[i]->has_observation();
    }
    for (int q = 0; q < m; ++q) {
        auto *lm = local_landmarks[q];
        valid[q] = lm->is_observable_in_tracking_ && !lm->will_be_erased();
        qx[q] = lm->reproj_in_tracking_(0);
        qy[q] = lm->reproj_in_tracking_(1);
        qxr[q] = lm->x_right_in_tracking_;
        lvl[q] = lm->scale_level_in_tracking_;
        const cv::Mat d = lm->get_descriptor();
        std::copy(d.data, d.data + 32, qdesc.begin() + (size_t)q * 32);
    }
    plp_frame_points fp{n, x.data(), y.data(), oct.data(), nullptr, xr.data(), frm.descriptors_.data, claimed.data()};
    plp_landmark_queries lq{m, qx.data(), qy.data(), qxr.data(), lvl.data(), qdesc.data(), valid.data()};
    const plp_grid g = grid_of(frm.camera_);
    uint32_t num = 0;
    check(plp_match_frame_and_landmarks(thread_ctx(), &fp, &g, frm.scale_factors_.data(), (int)frm.scale_factors_.size(),
                                        &lq, margin, lowe_ratio, best.data(), &num));
    for (int q = 0; q < m; ++q)  // re-apply the pointer writes in landmark order
        if (best[q] >= 0) frm.landmarks_[best[q]] = local_landmarks[q];
    return num;
}

// ---- optimize::pose_optimizer::optimize (optimize/pose_optimizer.cc:53-229) ----------------------------------
inline unsigned pose_optimize(PLPSLAM::data::frame &frm, bool with_lines, int num_trials = 4, int num_each_iter = 10) {
    std::vector<plp_pt_obs> pts;
    std::vector<unsigned> pt_idx;
    for (unsigned idx = 0; idx < frm.num_keypts_; ++idx) {
        auto lm = frm.landmarks_[idx];
        if (!lm || lm->will_be_erased()) continue;
        frm.outlier_flags_[idx] = false;
        const PLPSLAM::Vec3_t X = lm->get_pos_in_world();
        const auto &kp = frm.undist_keypts_[idx];
        pts.push_back(plp_pt_obs{{X(0), X(1), X(2)}, kp.pt.x, kp.pt.y, frm.stereo_x_right_[idx],
                                 frm.inv_level_sigma_sq_[kp.octave]});
        pt_idx.push_back(idx);
    }
    std::vector<plp_line_obs> lines;
    std::vector<unsigned> line_idx;
    if (with_lines && pts.size() >= 5) {
        for (unsigned idx = 0; idx < frm._num_keylines; ++idx) {
            auto ll = frm._landmarks_line[idx];
            if (!ll || ll->will_be_erased()) continue;
            frm._outlier_flags_line[idx] = false;
            const PLPSLAM::Vec6_t L = ll->get_PlueckerCoord();
            const auto &kl = frm._keylsd[idx];
            lines.push_back(plp_line_obs{{L(0), L(1), L(2), L(3), L(4), L(5)}, kl.getStartPoint().x, kl.getStartPoint().y,
                                         kl.getEndPoint().x, kl.getEndPoint().y, frm._inv_level_sigma_sq_lsd[kl.octave], 0.f});
            line_idx.push_back(idx);
        }
    }
    const plp_camera cam = camera_of(frm.camera_);
    const plp_pose_opt_cfg cfg{num_trials, num_each_iter};
    double T_in[16], T_out[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) T_in[r * 4 + c] = frm.cam_pose_cw_(r, c);
    std::vector<uint8_t> pout(pts.size() + 1), lout(lines.size() + 1);
    int32_t n_inliers = 0;
    check(plp_pose_optimize(thread_ctx(), &cam, T_in, pts.data(), (int)pts.size(), lines.data(), (int)lines.size(), &cfg,
                            T_out, pout.data(), lout.data(), &n_inliers));
    if (pts.size() < 5) return 0;  // pose_optimizer.cc:153-156
    for (size_t k = 0; k < pt_idx.size(); ++k) frm.outlier_flags_[pt_idx[k]] = pout[k];
    for (size_t k = 0; k < line_idx.size(); ++k) frm._outlier_flags_line[line_idx[k]] = lout[k];
    PLPSLAM::Mat44_t T;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) T(r, c) = T_out[r * 4 + c];
    frm.set_cam_pose(T);
    return (unsigned)n_inliers;
}

}  // namespace plpslam_b200

#endif  // PLPSLAM_B200_WITH_REFERENCE_TYPES
