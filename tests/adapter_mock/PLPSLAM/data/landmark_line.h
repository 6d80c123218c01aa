#pragma once
#include <opencv2/core.hpp>
#include "PLPSLAM/type.h"
namespace PLPSLAM { namespace data {
class keyframe;
class Line {  // data/landmark_line.h
public:
    Vec6_t get_pos_in_world() const; Vec6_t get_PlueckerCoord() const; cv::Mat get_descriptor() const;
    bool will_be_erased(); bool has_observation() const; bool is_observed_in_keyframe(keyframe *) const;
    float get_min_valid_distance() const; float get_max_valid_distance() const;
    unsigned predict_scale_level(const float &, const float &, const unsigned int &);
};
} }
