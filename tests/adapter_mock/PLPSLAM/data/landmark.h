#pragma once
#include <opencv2/core.hpp>
#include <map>
#include "PLPSLAM/type.h"
namespace plpslam_b200 { struct landmark_access; }
namespace PLPSLAM { namespace data {
class keyframe; class frame; class Plane;
class landmark {  // data/landmark.h
public:
    unsigned id_; std::map<keyframe *, unsigned> get_observations() const; void erase_observation(keyframe *);
    void set_pos_in_world(const Vec3_t &); void update_normal_and_depth();
    Vec3_t get_pos_in_world() const; Vec3_t get_obs_mean_normal() const; cv::Mat get_descriptor() const;
    bool will_be_erased(); bool has_observation() const; bool is_observed_in_keyframe(keyframe *) const;
    unsigned num_observations() const; void add_observation(keyframe *, unsigned); void replace(landmark *);
    float get_min_valid_distance() const; float get_max_valid_distance() const; Plane *get_Owning_Plane() const;
    unsigned predict_scale_level(const float, const frame *) const; unsigned predict_scale_level(const float, const keyframe *) const;
    Vec2_t reproj_in_tracking_; float x_right_in_tracking_; bool is_observable_in_tracking_; int scale_level_in_tracking_;
private:
    friend struct ::plpslam_b200::landmark_access;  // the one-line addition INTEGRATION.md asks for
    float min_valid_dist_ = 0, max_valid_dist_ = 0;
};
} }
