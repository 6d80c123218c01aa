#pragma once
#include "PLPSLAM/type.h"
namespace PLPSLAM { namespace camera {
enum class setup_type_t { Monocular = 0, Stereo = 1, RGBD = 2 };
struct image_bounds { float min_x_, max_x_, min_y_, max_y_; };
struct base {  // camera/base.h
    setup_type_t setup_type_; image_bounds img_bounds_; double inv_cell_width_, inv_cell_height_;
    unsigned num_grid_cols_, num_grid_rows_; double focal_x_baseline_, true_baseline_;
    virtual bool reproject_to_bearing(const Mat33_t &, const Vec3_t &, const Vec3_t &, Vec3_t &) const { return true; }
    virtual bool reproject_to_image(const Mat33_t &, const Vec3_t &, const Vec3_t &, Vec2_t &, float &) const { return true; }
    virtual ~base() {}
};
struct perspective : base { double fx_, fy_, cx_, cy_; };
} }
