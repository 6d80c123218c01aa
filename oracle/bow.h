/* oracle/bow.h -- DBoW2 vocabulary tree transform + match::bow_tree restatement (TEST INFRASTRUCTURE ONLY); see oracle.h.
 *
 * DBoW2 is an un-vendored, unversioned dependency of the reference (find_package(DBoW2), src/PLPSLAM/CMakeLists.txt; the
 * README points at OpenVSLAM's fork with loadFromBinaryFile).  Its arithmetic on this path is restated from the published
 * algorithm (TemplatedVocabulary::transform with levelsup, FORB::distance = 256-bit Hamming) and anchored on the
 * reference's call site (data/frame.cc:785-795: transform(descriptors, bow_vec_, bow_feat_vec_, 4)) and on the binary
 * vocabulary the reference ships (orb_vocab/orb_vocab.dbow2: header {n_nodes, node_size = 41, k, L, scoring, weighting},
 * then per node {int32 parent, 32 B descriptor, float weight, bool is_leaf}).  PARITY UNPINNED (no DBoW2 here). */
#ifndef PLP_ORACLE_BOW_H
#define PLP_ORACLE_BOW_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orc_bow_vocab orc_bow_vocab;
/* nodes 1 .. num_nodes-1 in file order (node 0 is the root); parent[i - 1] < i */
orc_bow_vocab *orc_bow_vocab_create(int k, int L, int num_nodes, const int32_t *parent, const uint8_t *desc,
                                    const float *weight, const uint8_t *is_leaf);
orc_bow_vocab *orc_bow_vocab_load(const char *path); /* loadFromBinaryFile (system.cc:82) */
void orc_bow_vocab_destroy(orc_bow_vocab *v);
void orc_bow_vocab_info(const orc_bow_vocab *v, int32_t *k, int32_t *L, int32_t *num_nodes, int32_t *num_words);
/* TemplatedVocabulary::transform(feature, word_id, weight, &nid, levelsup) for every row: tree descent by minimum Hamming
 * distance (first child wins ties), node_id = the node passed at level L - levelsup (0 = root when L <= levelsup).
 * weight 0 marks a stopped word (skipped by transform(features, v, fv, levelsup)). */
void orc_bow_transform(const orc_bow_vocab *v, const uint8_t *desc, int n, int levelsup, int32_t *word_id_out,
                       int32_t *node_id_out, float *weight_out);
#ifdef __cplusplus
}
#endif
#endif
