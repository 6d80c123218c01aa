#pragma once
#include <vector>
namespace PLPSLAM { namespace data {
class keyframe;
class graph_node {  // data/graph_node.h
public:
    std::vector<keyframe *> get_covisibilities() const;
};
} }
