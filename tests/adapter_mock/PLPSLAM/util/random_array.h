#pragma once
#include <cstddef>
#include <vector>
namespace PLPSLAM { namespace util { template <typename T> std::vector<T> create_random_array(const size_t size, const T rand_min, const T rand_max); } }
