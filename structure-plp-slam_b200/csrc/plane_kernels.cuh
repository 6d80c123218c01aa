// plane_kernels.cuh -- device code of the plane RANSAC (plane.cu launches it).  Kept free of host-side CUDA runtime
// dependencies so that tests/cta_emu can compile the SAME text for the host (threads + a barrier stand in for a CTA) and
// check the kernel logic against the oracle without a GPU.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/plpslam_b200.h"
#include "planemath.h"

namespace plp {

namespace {

constexpr int kPlThreads = 128;

struct PlaneJob {
    const double *pos;      // n x 3
    const uint8_t *valid;   // may be null
    const int32_t *samples; // num_iter x sample_size
    int n, num_iter, sample_size;
    plp_plane_ransac_cfg cfg;
    // per hypothesis
    double *eq_s, *eq_r;    // num_iter x 4: sample fit, refit on the inliers
    double *res, *err;      // num_iter: sample residual, refit error
    int32_t *elig, *cnt;    // num_iter: refit happened, inlier count
    uint8_t *flag;          // num_iter x n scratch
    int32_t *idx;           // num_iter x n: inlier indices in ascending order
    // state / result
    double *eq, *plane_err; // 4 / 1: in = the Plane before the call, out = after
    uint8_t *inlier;        // n
    int32_t *status;        // 1
};

__global__ void __launch_bounds__(kPlThreads) plane_hypothesis_kernel(PlaneJob J) {
    __shared__ double s_eq[4];
    const int it = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {  // [1] fit the sample (:460)
        double e[4];
        const PlaneSelIndices sel{J.samples + (size_t)it * J.sample_size, J.sample_size};
        J.res[it] = plane_fit(J.pos, sel, e);
        for (int k = 0; k < 4; ++k) {
            s_eq[k] = e[k];
            J.eq_s[it * 4 + k] = e[k];
        }
    }
    __syncthreads();
    uint8_t *flag = J.flag + (size_t)it * J.n;
    for (int j = tid; j < J.n; j += kPlThreads) {  // [2] :472-487
        const bool v = J.valid ? J.valid[j] != 0 : true;
        flag[j] = (v && plane_distance(s_eq, J.pos + 3 * (size_t)j) < J.cfg.planar_distance_thresh) ? 1 : 0;
    }
    __syncthreads();
    if (tid == 0) {  // [3] compact in index order, refit when eligible
        int32_t *idx = J.idx + (size_t)it * J.n;
        int c = 0;
        for (int j = 0; j < J.n; ++j)
            if (flag[j]) idx[c++] = j;
        J.cnt[it] = c;
        bool eligible;
        if (J.cfg.mode == 0) {
            const double inlier_ratio = (double)c / (double)J.n;
            eligible = inlier_ratio > J.cfg.inliers_ratio_thr && c >= J.cfg.points_per_ransac;  // :494-499
        } else {
            eligible = c >= J.cfg.points_per_ransac;  // :664
        }
        J.elig[it] = eligible ? 1 : 0;
        double e[4] = {0, 0, 0, 0}, error = 0.0;
        if (eligible) {
            const PlaneSelIndices isel{idx, c};
            error = plane_fit(J.pos, isel, e);
        }
        J.err[it] = error;
        for (int k = 0; k < 4; ++k) J.eq_r[it * 4 + k] = e[k];
    }
}

__global__ void __launch_bounds__(kPlThreads) plane_select_kernel(PlaneJob J) {
    __shared__ int s_best_it, s_ok, s_kept;
    __shared__ double s_eq[4];
    const int tid = threadIdx.x;
    for (int j = tid; j < J.n; j += kPlThreads) J.inlier[j] = 0;
    if (tid == 0) {
        double best_error = J.cfg.mode == 1 ? J.cfg.initial_best_error : 1.7976931348623157e308;  // :439 / :609
        bool found = false;
        int best_it = -1;
        double eq[4] = {J.eq[0], J.eq[1], J.eq[2], J.eq[3]}, plane_err = *J.plane_err;
        for (int it = 0; it < J.num_iter; ++it) {
            const double residual = J.res[it];
            if (residual < best_error) best_error = residual;      // :461-464
            for (int k = 0; k < 4; ++k) eq[k] = J.eq_s[it * 4 + k];  // plane->set_equation(sample fit), every iteration
            plane_err = residual;                                   // plane->set_best_error(residual)
            if (J.elig[it]) {
                const double error = J.err[it];
                if (error < best_error) {  // :505 / :669
                    best_error = error;
                    for (int k = 0; k < 4; ++k) eq[k] = J.eq_r[it * 4 + k];
                    plane_err = best_error;
                    best_it = it;
                    found = true;
                    if (J.cfg.mode == 0 && error < J.cfg.final_error_thresh) break;  // :526-534
                }
            }
        }
        for (int k = 0; k < 4; ++k) {
            J.eq[k] = eq[k];
            s_eq[k] = eq[k];
        }
        *J.plane_err = plane_err;
        s_best_it = best_it;
        s_ok = (found && !(best_error > J.cfg.final_error_thresh)) ? 1 : 0;  // :545-562 / :691
        s_kept = 0;
    }
    __syncthreads();
    if (!s_ok) {
        if (tid == 0) *J.status = 0;
        return;
    }
    // [4] :565-578 / :698-711: the equation the Plane holds NOW filters the best inlier list
    const int32_t *idx = J.idx + (size_t)s_best_it * J.n;
    const int c = J.cnt[s_best_it];
    int local = 0;
    for (int k = tid; k < c; k += kPlThreads) {
        const int j = idx[k];
        if (plane_distance(s_eq, J.pos + 3 * (size_t)j) < J.cfg.planar_distance_thresh) {
            J.inlier[j] = 1;
            ++local;
        }
    }
    atomicAdd(&s_kept, local);
    __syncthreads();
    if (J.cfg.mode == 1 && s_kept < J.cfg.points_per_ransac) {  // :713-717
        for (int k = tid; k < c; k += kPlThreads) J.inlier[idx[k]] = 0;
        if (tid == 0) *J.status = 2;
        return;
    }
    if (tid == 0) *J.status = 1;
}

}  // namespace

}  // namespace plp
