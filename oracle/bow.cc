// oracle/bow.cc -- DBoW2 vocabulary tree (binary format + transform), TEST INFRASTRUCTURE ONLY.  See bow.h.
#include "bow.h"
#include "oracle.h"

#include <cstdio>
#include <cstring>
#include <vector>

struct orc_bow_vocab {
    int k = 0, L = 0;
    struct Node {
        int32_t parent = 0;
        std::vector<int32_t> children;  // in file order (m_nodes[parent].children.push_back(n_id))
        uint8_t desc[32] = {0};
        float weight = 0.f;
        int32_t word_id = -1;
        bool leaf = false;
    };
    std::vector<Node> nodes;
    int num_words = 0;
};

extern "C" {

orc_bow_vocab *orc_bow_vocab_create(int k, int L, int num_nodes, const int32_t *parent, const uint8_t *desc,
                                    const float *weight, const uint8_t *is_leaf) {
    if (k < 1 || L < 1 || num_nodes < 1) return nullptr;
    auto *v = new orc_bow_vocab;
    v->k = k;
    v->L = L;
    v->nodes.resize(num_nodes);
    for (int id = 1; id < num_nodes; ++id) {
        auto &n = v->nodes[id];
        n.parent = parent[id - 1];
        if (n.parent < 0 || n.parent >= id) {
            delete v;
            return nullptr;
        }
        v->nodes[n.parent].children.push_back(id);
        std::memcpy(n.desc, desc + 32 * (size_t)(id - 1), 32);
        n.weight = weight[id - 1];
        n.leaf = is_leaf[id - 1] != 0;
        if (n.leaf) n.word_id = v->num_words++;  // words are numbered in file order
    }
    return v;
}

orc_bow_vocab *orc_bow_vocab_load(const char *path) {
    FILE *f = std::fopen(path, "rb");
    if (!f) return nullptr;
    uint32_t n_nodes = 0, node_size = 0;
    int32_t k = 0, L = 0, scoring = 0, weighting = 0;
    bool ok = std::fread(&n_nodes, 4, 1, f) == 1 && std::fread(&node_size, 4, 1, f) == 1 && std::fread(&k, 4, 1, f) == 1 &&
              std::fread(&L, 4, 1, f) == 1 && std::fread(&scoring, 4, 1, f) == 1 && std::fread(&weighting, 4, 1, f) == 1;
    if (!ok || node_size != 41 || n_nodes < 1) {
        std::fclose(f);
        return nullptr;
    }
    const size_t cnt = n_nodes - 1;
    std::vector<uint8_t> raw(cnt * 41);
    ok = std::fread(raw.data(), 41, cnt, f) == cnt;
    std::fclose(f);
    if (!ok) return nullptr;
    std::vector<int32_t> parent(cnt);
    std::vector<uint8_t> desc(cnt * 32), leaf(cnt);
    std::vector<float> weight(cnt);
    for (size_t i = 0; i < cnt; ++i) {
        const uint8_t *r = raw.data() + 41 * i;
        std::memcpy(&parent[i], r, 4);
        std::memcpy(&desc[32 * i], r + 4, 32);
        std::memcpy(&weight[i], r + 36, 4);
        leaf[i] = r[40];
    }
    return orc_bow_vocab_create(k, L, (int)n_nodes, parent.data(), desc.data(), weight.data(), leaf.data());
}

void orc_bow_vocab_destroy(orc_bow_vocab *v) { delete v; }

void orc_bow_vocab_info(const orc_bow_vocab *v, int32_t *k, int32_t *L, int32_t *num_nodes, int32_t *num_words) {
    *k = v->k;
    *L = v->L;
    *num_nodes = (int32_t)v->nodes.size();
    *num_words = v->num_words;
}

void orc_bow_transform(const orc_bow_vocab *v, const uint8_t *desc, int n, int levelsup, int32_t *word_id_out,
                       int32_t *node_id_out, float *weight_out) {
    const int nid_level = v->L - levelsup;
    for (int i = 0; i < n; ++i) {
        const uint8_t *f = desc + 32 * (size_t)i;
        int32_t nid = 0;  // "if (nid_level <= 0) *nid = 0"; a leaf above nid_level leaves it at the root as well
        int32_t final_id = 0;
        int current_level = 0;
        if (v->nodes[0].children.empty()) {  // empty vocabulary: transform() returns without touching anything
            word_id_out[i] = -1;
            node_id_out[i] = 0;
            weight_out[i] = 0.f;
            continue;
        }
        do {
            ++current_level;
            const auto &nodes = v->nodes[final_id].children;
            final_id = nodes[0];
            unsigned best_d = orc_hamming_32(f, v->nodes[final_id].desc);
            for (size_t c = 1; c < nodes.size(); ++c) {
                const unsigned d = orc_hamming_32(f, v->nodes[nodes[c]].desc);
                if (d < best_d) {
                    best_d = d;
                    final_id = nodes[c];
                }
            }
            if (current_level == nid_level) nid = final_id;
        } while (!v->nodes[final_id].children.empty());
        word_id_out[i] = v->nodes[final_id].word_id;
        weight_out[i] = v->nodes[final_id].weight;
        node_id_out[i] = nid;
    }
}

}  // extern "C"
