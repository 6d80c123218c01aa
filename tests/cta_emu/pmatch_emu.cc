// pmatch_emu.cc -- csrc/point_match_kernels.cuh (the window matcher of projection::match_frame_and_landmarks /
// match_current_and_last_frames) executed on the host: 1024 threads, counting sort with shared-memory atomics, 4-lane
// query groups merged with shuffles, deferred acceptance with atomicMin.
#include "cta_emu.h"

#include <string.h>

#include "point_match_kernels.cuh"

using namespace plp;

extern "C" void emu_point_match(const plp_grid *grid, int n, const float *x, const float *y, const int32_t *octave,
                                const float *angle, const float *x_right, const uint8_t *desc, const uint8_t *claimed, int m,
                                const float *qx, const float *qy, const float *qxr, const float *qradius, const int32_t *qmin,
                                const int32_t *qmax, const float *qangle, const uint8_t *qdesc, const uint8_t *qvalid,
                                unsigned hamm_thr_p1, int ratio_test, float lowe_ratio, int check_orientation, int cap,
                                int32_t *choice_scratch, int32_t *best_idx_out, int32_t *matched_out, uint32_t *num_matches) {
    PointMatchJob J;
    memset(&J, 0, sizeof(J));
    J.n = n;
    J.x = x;
    J.y = y;
    J.octave = octave;
    J.angle = angle;
    J.x_right = x_right;
    J.desc = desc;
    J.claimed = claimed;
    J.m = m;
    J.qx = qx;
    J.qy = qy;
    J.qxr = qxr;
    J.qradius = qradius;
    J.qmin = qmin;
    J.qmax = qmax;
    J.qangle = qangle;
    J.qdesc = qdesc;
    J.qvalid = qvalid;
    J.choice = choice_scratch;
    J.best_idx_out = best_idx_out;
    J.matched_out = matched_out;
    J.num_matches = num_matches;
    J.hamm_thr_p1 = hamm_thr_p1;
    const PointMatchJob *jobs = &J;
    const size_t smem = pm::point_smem_bytes(cap, grid->num_cols, grid->num_rows);
    emu_launch2(pm::point_match_kernel, 1u, 1u, (unsigned)pm::kThreads, smem, jobs, *grid, cap, ratio_test, lowe_ratio,
                check_orientation);
}
