#pragma once
#include <map>
#include <vector>
namespace DBoW2 {
typedef unsigned int WordId; typedef double WordValue; typedef unsigned int NodeId;
enum LNorm { L1, L2 };
struct BowVector : std::map<WordId, WordValue> { void addWeight(WordId, WordValue) {} void normalize(LNorm) {} };
struct FeatureVector : std::map<NodeId, std::vector<unsigned int>> { void addFeature(NodeId, unsigned int) {} };
}
