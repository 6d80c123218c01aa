#pragma once
#include <memory>
#include "PLPSLAM/data/frame.h"
#include "PLPSLAM/data/graph_node.h"
namespace PLPSLAM { namespace data {
class keyframe {  // data/keyframe.h
public:
    unsigned num_keypts_; std::vector<cv::KeyPoint> keypts_, undist_keypts_; std::vector<float> stereo_x_right_;
    eigen_alloc_vector<Vec3_t> bearings_; cv::Mat descriptors_; camera::base *camera_;
    std::vector<float> scale_factors_, inv_level_sigma_sq_; unsigned num_scale_levels_; float log_scale_factor_;
    DBoW2::BowVector bow_vec_; DBoW2::FeatureVector bow_feat_vec_;
    std::vector<cv::line_descriptor::KeyLine> _keylsd; cv::Mat _lbd_descr;
    unsigned id_; std::unique_ptr<graph_node> graph_node_; std::vector<float> _inv_level_sigma_sq_lsd;
    bool will_be_erased(); void set_cam_pose(const Mat44_t &); float compute_median_depth(bool abs) const;
    std::vector<Line *> get_landmarks_line() const; void erase_landmark(landmark *); void erase_landmark_line(Line *);
    std::vector<landmark *> get_landmarks() const; landmark *get_landmark(unsigned) const; void add_landmark(landmark *, unsigned);
    Mat33_t get_rotation() const; Vec3_t get_translation() const; Vec3_t get_cam_center() const; Mat44_t get_cam_pose() const;
};
} }
