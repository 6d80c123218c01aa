#pragma once
#include <vector>
#include "PLPSLAM/data/landmark.h"
namespace PLPSLAM { namespace data {
class Plane {  // data/landmark_plane.h:59-84
public:
    std::vector<landmark *> get_landmarks() const; void set_landmarks(std::vector<landmark *> &lms);
    void set_equation(double a, double b, double c, double d); void get_equation(double &a, double &b, double &c, double &d) const;
    void set_invalid(); void set_need_refinement(); void set_landmarks_ownership(); void remove_landmarks_ownership();
    void set_best_error(double const &error); double get_best_error() const;
    unsigned _id; bool is_valid() const; bool need_refinement() const; Vec3_t get_normal() const; double get_offset() const;
};
} }
