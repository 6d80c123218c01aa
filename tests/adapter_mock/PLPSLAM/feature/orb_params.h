#pragma once
namespace PLPSLAM { namespace feature { struct orb_params { unsigned max_num_keypts_; float scale_factor_; unsigned num_levels_; unsigned ini_fast_thr_; unsigned min_fast_thr; }; } }
