// plane_emu.cc -- csrc/plane_kernels.cuh executed on the host (see cta_emu.h).  The host wrapper below repeats the
// argument handling of plp_plane_ransac (csrc/plane.cu) around the two kernel launches.
#include "cta_emu.h"

#include <string.h>

#include "plane_kernels.cuh"

using namespace plp;

extern "C" int emu_plane_ransac(const double *pos_w, const uint8_t *valid, int n, const int32_t *samples, int num_iter,
                                int sample_size, const plp_plane_ransac_cfg *cfg, double *eq_inout,
                                double *plane_error_inout, uint8_t *inlier_out) {
    if (n > 0) memset(inlier_out, 0, (size_t)n);
    if (n == 0) return 0;
    if (n < cfg->points_per_ransac) return cfg->mode == 1 ? 2 : 0;
    if (num_iter == 0) return 0;
    const size_t N = (size_t)n, K = (size_t)num_iter;
    std::vector<double> eq_s(K * 4), eq_r(K * 4), res(K), err(K);
    std::vector<int32_t> elig(K), cnt(K), idx(K * N);
    std::vector<uint8_t> flag(K * N);
    int32_t status = -1;
    PlaneJob J;
    J.pos = pos_w;
    J.valid = valid;
    J.samples = samples;
    J.n = n;
    J.num_iter = num_iter;
    J.sample_size = sample_size;
    J.cfg = *cfg;
    J.eq_s = eq_s.data();
    J.eq_r = eq_r.data();
    J.res = res.data();
    J.err = err.data();
    J.elig = elig.data();
    J.cnt = cnt.data();
    J.flag = flag.data();
    J.idx = idx.data();
    J.eq = eq_inout;
    J.plane_err = plane_error_inout;
    J.inlier = inlier_out;
    J.status = &status;
    emu_launch(plane_hypothesis_kernel, (unsigned)num_iter, (unsigned)kPlThreads, J);
    emu_launch(plane_select_kernel, 1u, (unsigned)kPlThreads, J);
    return status;
}
