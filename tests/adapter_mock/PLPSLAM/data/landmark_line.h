#pragma once
#include <opencv2/core.hpp>
#include <map>
#include "PLPSLAM/type.h"
namespace PLPSLAM { namespace data {
class keyframe;
class Line {  // data/landmark_line.h
public:
    unsigned _id; std::map<keyframe *, unsigned> get_observations() const; void erase_observation(keyframe *);
    keyframe *get_ref_keyframe() const; int get_index_in_keyframe(keyframe *) const;
    void set_PlueckerCoord_without_update_endpoints(const Vec6_t &); void set_pos_in_world_without_update_pluecker(const Vec6_t &);
    void update_information(); void prepare_for_erasing();
    Vec2_t _reproj_in_tracking_sp, _reproj_in_tracking_ep; bool _is_observable_in_tracking; int _scale_level_in_tracking;
    Vec6_t get_pos_in_world() const; Vec6_t get_PlueckerCoord() const; cv::Mat get_descriptor() const;
    bool will_be_erased(); bool has_observation() const; bool is_observed_in_keyframe(keyframe *) const;
    float get_min_valid_distance() const; float get_max_valid_distance() const;
    unsigned predict_scale_level(const float &, const float &, const unsigned int &);
};
} }
