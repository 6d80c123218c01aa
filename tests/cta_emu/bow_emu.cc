// bow_emu.cc -- csrc/bow_kernels.cuh (DBoW2 transform + match::bow_tree, both GPU-verified) executed on the host.
#include "cta_emu.h"

#include <string.h>

#include "bow_kernels.cuh"

using namespace plp;

extern "C" void emu_bow_transform(int G, const uint8_t *node_desc, const uint32_t *child_begin, const uint32_t *children,
                                  const float *weight, const int32_t *word_id, const uint8_t *desc, int n, int nid_level,
                                  int32_t *word_out, int32_t *node_out, float *weight_out) {
    VocabDev V;
    V.desc = node_desc;
    V.child_begin = child_begin;
    V.children = children;
    V.weight = weight;
    V.word_id = word_id;
    const unsigned blocks = (unsigned)((n + 256 / G - 1) / (256 / G));
    blockDim.x = 256;
    if (G == 4) emu_launch(bow_transform_kernel<4>, blocks, 256u, V, desc, n, nid_level, word_out, node_out, weight_out);
    else if (G == 8) emu_launch(bow_transform_kernel<8>, blocks, 256u, V, desc, n, nid_level, word_out, node_out, weight_out);
    else if (G == 16) emu_launch(bow_transform_kernel<16>, blocks, 256u, V, desc, n, nid_level, word_out, node_out, weight_out);
    else emu_launch(bow_transform_kernel<32>, blocks, 256u, V, desc, n, nid_level, word_out, node_out, weight_out);
}

extern "C" unsigned emu_bow_match(int n1, const uint8_t *desc1, const float *angle1, const uint8_t *valid1, int n2,
                                  const uint8_t *desc2, const float *angle2, const uint8_t *valid2, const uint32_t *idx1,
                                  const uint32_t *idx2, int num_nodes, const int32_t *nb1, const int32_t *ne1,
                                  const int32_t *nb2, const int32_t *ne2, float lowe_ratio, int check_orientation,
                                  int32_t *m21, int32_t *m12) {
    std::vector<uint8_t> claimed((size_t)n2 + 1);
    std::vector<int32_t> choice((size_t)n1 + 1);
    uint32_t num = 0;
    BowJob J;
    memset(&J, 0, sizeof(J));
    J.n1 = n1;
    J.n2 = n2;
    J.num_nodes = num_nodes;
    J.desc1 = desc1;
    J.desc2 = desc2;
    J.angle1 = angle1;
    J.angle2 = angle2;
    J.valid1 = valid1;
    J.valid2 = valid2;
    J.idx1 = idx1;
    J.idx2 = idx2;
    J.nb1 = nb1;
    J.ne1 = ne1;
    J.nb2 = nb2;
    J.ne2 = ne2;
    J.claimed = claimed.data();
    J.choice = choice.data();
    J.matched_2_of_1 = m21;
    J.matched_1_of_2 = m12;
    J.num_matches = &num;
    const BowJob *jobs = &J;
    emu_launch(bow_match_kernel, 1u, (unsigned)kMatchThreads, jobs, lowe_ratio, check_orientation);
    return num;
}
