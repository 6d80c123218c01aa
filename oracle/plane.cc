// oracle/plane.cc -- Planar_Mapping_module::estimate_plane_sequential_RANSAC / update_plane_via_RANSAC / estimate_plane_SVD
// (TEST INFRASTRUCTURE ONLY).  Follows /root/reference/src/PLPSLAM/planar_mapping_module.cc:412-771.
#include "plane.h"
#include "planemath.h"

#include <limits>
#include <vector>

extern "C" {

double orc_plane_fit(const double *pos_w, const int32_t *idx, int cnt, double *eq_out) {
    const PlaneSelIndices sel{idx, cnt};
    return plane_fit(pos_w, sel, eq_out);
}

int orc_plane_ransac(const double *pos_w, const uint8_t *valid, int n, const int32_t *samples, int num_iter, int sample_size,
                     const orc_plane_cfg *cfg, double *eq_out, double *plane_error_out, uint8_t *inlier_out) {
    const int P = cfg->points_per_ransac;
    for (int j = 0; j < n; ++j) inlier_out[j] = 0;
    if (n == 0) return 0;                            // :423-426 / :597-600
    if (n < P) return cfg->mode == 1 ? 2 : 0;        // :428-436 / :602-606
    double best_error = cfg->mode == 1 ? cfg->initial_best_error : std::numeric_limits<double>::max();  // :439 / :609
    bool best_found = false;
    std::vector<int> inliers_list, best_inliers_list;
    double eq[4] = {eq_out[0], eq_out[1], eq_out[2], eq_out[3]}, plane_error = *plane_error_out;  // the Plane's state
    for (int i = 0; i < num_iter; ++i) {
        // [1] :460-470 / :634-644
        double e[4];
        const PlaneSelIndices sel{samples + (size_t)i * sample_size, sample_size};
        const double residual = plane_fit(pos_w, sel, e);
        if (residual < best_error) best_error = residual;
        for (int k = 0; k < 4; ++k) eq[k] = e[k];  // plane->set_equation(a_best, ...)
        plane_error = residual;                     // plane->set_best_error(residual)
        // [2] :472-487 / :646-661
        inliers_list.clear();
        for (int j = 0; j < n; ++j) {
            if (valid && !valid[j]) continue;
            if (plane_distance(eq, pos_w + 3 * (size_t)j) < cfg->planar_distance_thresh) inliers_list.push_back(j);
        }
        // [3]
        bool eligible;
        if (cfg->mode == 0) {
            const double inlier_ratio = double(inliers_list.size()) / double(n);
            eligible = inlier_ratio > cfg->inliers_ratio_thr && (int)inliers_list.size() >= P;  // :494-499
        } else {
            eligible = (int)inliers_list.size() >= P;  // :664
        }
        if (eligible) {
            const PlaneSelIndices isel{inliers_list.data(), (int)inliers_list.size()};
            const double error = plane_fit(pos_w, isel, e);
            if (error < best_error) {  // :505 / :669
                best_error = error;
                for (int k = 0; k < 4; ++k) eq[k] = e[k];
                plane_error = best_error;
                best_inliers_list = inliers_list;
                best_found = true;
                if (cfg->mode == 0 && error < cfg->final_error_thresh) break;  // :526-534 (estimate only)
            }
        }
    }
    for (int k = 0; k < 4; ++k) eq_out[k] = eq[k];
    *plane_error_out = plane_error;
    if (!best_found) return 0;                             // :545-553 / :691
    if (best_error > cfg->final_error_thresh) return 0;    // :554-562 / :691
    // [4] :565-578 / :698-711: the CURRENT plane equation (the last one set) filters the best inlier list
    int kept = 0;
    for (const int j : best_inliers_list) {
        if (valid && !valid[j]) continue;
        if (plane_distance(eq, pos_w + 3 * (size_t)j) < cfg->planar_distance_thresh) {
            inlier_out[j] = 1;
            ++kept;
        }
    }
    if (cfg->mode == 1 && kept < P) {  // :713-717
        for (int j = 0; j < n; ++j) inlier_out[j] = 0;
        return 2;
    }
    return 1;
}

}  // extern "C"
