// essential_emu.cc -- csrc/essential_kernels.cuh executed on the host (see cta_emu.h).  These kernels are GPU-verified
// (tests/test_essential_gpu.py); running them here as well validates the emulator itself.
#include "cta_emu.h"

#include <string.h>

#include "essential_kernels.cuh"

using namespace plp;

extern "C" int emu_essential_ransac(const double *b1, const double *b2, const int32_t *matches, int num_matches,
                                    const int32_t *samples, int num_iter, int recompute, uint8_t *is_inlier_out,
                                    double *best_E_out, double *best_score_out) {
    if (num_matches < 8) return 0;
    const size_t M = (size_t)num_matches, K = (size_t)num_iter;
    std::vector<double> E(K * 9);
    std::vector<float> score(K), res(K * M * 2);
    std::vector<uint8_t> inl(K * M);
    int32_t valid = 0;
    EssJob J;
    J.b1 = b1;
    J.b2 = b2;
    J.matches = matches;
    J.samples = samples;
    J.num_matches = num_matches;
    J.num_iter = num_iter;
    J.recompute = recompute;
    J.E = E.data();
    J.score = score.data();
    J.inlier = inl.data();
    J.res = res.data();
    J.best_inlier = is_inlier_out;
    J.best_E = best_E_out;
    J.best_score = best_score_out;
    J.valid = &valid;
    emu_launch(essential_hypothesis_kernel, (unsigned)num_iter, (unsigned)kEssThreads, J);
    emu_launch(essential_select_kernel, 1u, (unsigned)kEssThreads, J);
    return valid;
}
