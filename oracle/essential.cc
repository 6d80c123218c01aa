// oracle/essential.cc -- solve::essential_solver (TEST INFRASTRUCTURE ONLY).  Follows
// /root/reference/src/PLPSLAM/solve/essential_solver.cc; the SVDs are restated in essmath.h (see its header).
#include "essential.h"
#include "essmath.h"

#include <cstring>
#include <vector>

namespace {

// essential_solver.cc:200-254
float check_inliers(const double *E21, const double *b1, const double *b2, const int32_t *matches, int num_matches,
                    uint8_t *is_inlier) {
    float score = 0;
    for (int i = 0; i < num_matches; ++i) {
        float s2, s1;
        int add1;
        is_inlier[i] = (uint8_t)ess_check_match(E21, b1 + 3 * (size_t)matches[2 * i], b2 + 3 * (size_t)matches[2 * i + 1], &s2,
                                                &add1, &s1);
        score += s2;  // 0 when the first test failed (the reference `continue`s before adding)
        if (add1) score += s1;
    }
    return score;
}

}  // namespace

extern "C" {

void orc_essential_compute_E21(const double *bearings_1, const double *bearings_2, int n, double *E_21_out) {
    double ata[81];
    for (double &x : ata) x = 0.0;
    for (int i = 0; i < n; ++i) ess_accumulate(ata, bearings_1 + 3 * (size_t)i, bearings_2 + 3 * (size_t)i);
    ess_solve(ata, E_21_out);
}

int orc_essential_ransac(const double *b1, const double *b2, const int32_t *matches, int num_matches,
                         const int32_t *samples, int num_iter, int recompute, uint8_t *is_inlier_out,
                         double *best_E_21_out, double *best_score_out, float *scores_out) {
    constexpr int min_set_size = 8;
    for (int i = 0; i < num_matches; ++i) is_inlier_out[i] = 0;
    for (int k = 0; k < 9; ++k) best_E_21_out[k] = 0.0;
    *best_score_out = 0.0;
    if (num_matches < min_set_size) return 0;  // :45-49
    double best_score = 0.0;
    std::vector<uint8_t> in_sac(num_matches);
    for (int iter = 0; iter < num_iter; ++iter) {
        double ata[81], E[9];
        for (double &x : ata) x = 0.0;
        for (int i = 0; i < min_set_size; ++i) {  // :72-78
            const int idx = samples[iter * min_set_size + i];
            ess_accumulate(ata, b1 + 3 * (size_t)matches[2 * idx], b2 + 3 * (size_t)matches[2 * idx + 1]);
        }
        ess_solve(ata, E);                                                      // :81
        const float score = check_inliers(E, b1, b2, matches, num_matches, in_sac.data());  // :84
        if (scores_out) scores_out[iter] = score;
        if (best_score < score) {  // :87-92
            best_score = score;
            std::memcpy(best_E_21_out, E, sizeof(E));
            std::memcpy(is_inlier_out, in_sac.data(), num_matches);
        }
    }
    int num_inliers = 0;
    for (int i = 0; i < num_matches; ++i) num_inliers += is_inlier_out[i];
    const int valid = (best_score > 0.0) && (num_inliers >= min_set_size);  // :95-96
    *best_score_out = best_score;
    if (!recompute || !valid) return valid;
    // :103-120 recompute with all inliers
    double ata[81];
    for (double &x : ata) x = 0.0;
    for (int i = 0; i < num_matches; ++i)
        if (is_inlier_out[i]) ess_accumulate(ata, b1 + 3 * (size_t)matches[2 * i], b2 + 3 * (size_t)matches[2 * i + 1]);
    ess_solve(ata, best_E_21_out);
    *best_score_out = check_inliers(best_E_21_out, b1, b2, matches, num_matches, is_inlier_out);
    return valid;
}

}  // extern "C"
