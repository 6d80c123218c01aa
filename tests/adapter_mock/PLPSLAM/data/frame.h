#pragma once
#include <vector>
#include <opencv2/core.hpp>
#include <DBoW2/BowVector.h>
#include "PLPSLAM/type.h"
#include "PLPSLAM/camera/perspective.h"
#include "PLPSLAM/feature/orb_params.h"
#include "PLPSLAM/util/random_array.h"
namespace PLPSLAM { class system; namespace data {
class landmark; class Line;
class frame {  // data/frame.h
public:
    unsigned num_keypts_; std::vector<cv::KeyPoint> keypts_, undist_keypts_; std::vector<float> stereo_x_right_, depths_;
    eigen_alloc_vector<Vec3_t> bearings_; cv::Mat descriptors_; std::vector<landmark *> landmarks_; std::vector<bool> outlier_flags_;
    camera::base *camera_; std::vector<float> scale_factors_, inv_scale_factors_, level_sigma_sq_, inv_level_sigma_sq_;
    unsigned num_scale_levels_; float scale_factor_, log_scale_factor_;
    DBoW2::BowVector bow_vec_; DBoW2::FeatureVector bow_feat_vec_;
    unsigned _num_keylines; std::vector<cv::line_descriptor::KeyLine> _keylsd; cv::Mat _lbd_descr; std::vector<Line *> _landmarks_line;
    std::vector<bool> _outlier_flags_line; std::vector<float> _scale_factors_lsd, _inv_level_sigma_sq_lsd;
    Mat44_t cam_pose_cw_; void set_cam_pose(const Mat44_t &);
};
} }
