// poseopt_emu.cc -- csrc/pose_opt_kernels.cuh (motion-only BA, one 128-thread CTA per frame) executed on the host.
#include "cta_emu.h"

#include <string.h>

#include <vector>

static inline float sqrtf_(float v) { return sqrtf(v); }
#include "pose_opt_kernels.cuh"

using namespace plp;

extern "C" void emu_pose_optimize_batch(const plp_camera *cam, int batch, const double *T_in, const plp_pt_obs *pts,
                                        const int32_t *pt_off, const plp_line_obs *lines, const int32_t *line_off,
                                        int num_trials, int num_each_iter, double *T_out, uint8_t *pt_outlier,
                                        uint8_t *line_outlier, int32_t *n_inliers, int32_t *lm_iters) {
    std::vector<PoseJob> jobs(batch);
    for (int b = 0; b < batch; ++b) {
        PoseJob &J = jobs[b];
        J.T_in = T_in + 16 * (size_t)b;
        J.pts = pts + pt_off[b];
        J.n_pts = pt_off[b + 1] - pt_off[b];
        const int l0 = line_off ? line_off[b] : 0, l1 = line_off ? line_off[b + 1] : 0;
        J.lines = lines ? lines + l0 : nullptr;
        J.n_lines = lines ? l1 - l0 : 0;
        J.T_out = T_out + 16 * (size_t)b;
        J.pt_outlier = pt_outlier + pt_off[b];
        J.line_outlier = line_outlier ? line_outlier + l0 : nullptr;
        J.n_inliers = n_inliers + b;
        J.lm_iters = lm_iters + b;
    }
    const PoseJob *d_jobs = jobs.data();
    plp_pose_opt_cfg cfg{num_trials, num_each_iter};
    emu_launch(po::pose_opt_kernel, (unsigned)batch, (unsigned)po::kThreads,
               d_jobs, batch, *cam, cfg);
}
