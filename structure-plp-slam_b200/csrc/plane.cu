// plane.cu -- Planar_Mapping_module plane RANSAC (planar_mapping_module.cc:412-771) (sm_100a).
//
// Same shape as essential.cu: hypotheses are independent given their index draws -> one CTA per hypothesis (thread 0
// fits the sample with planemath.h -- the text the oracle compiles, hence bit-identical --, all threads test the
// landmarks, thread 0 compacts the inliers in index order and refits), then a one-CTA kernel replays the reference's
// sequential bookkeeping over the per-hypothesis results and applies step [4].  FP64, compiled with -fmad=false.
#include "common.cuh"
#include "pack.cuh"
#include "plane_kernels.cuh"


using namespace plp;

extern "C" {

plp_status plp_plane_ransac(plp_ctx *ctx, const double *pos_w, const uint8_t *valid, int n, const int32_t *samples,
                            int num_iter, int sample_size, const plp_plane_ransac_cfg *cfg, double *eq_inout,
                            double *plane_error_inout, uint8_t *inlier_out, int32_t *status_out) {
    PLP_REQUIRE(ctx && cfg && status_out, "null pointer");
    PLP_REQUIRE(n >= 0 && num_iter >= 0 && sample_size >= 0, "sizes");
    PLP_REQUIRE(cfg->mode == 0 || cfg->mode == 1, "mode");
    PLP_REQUIRE(cfg->points_per_ransac >= 1, "points_per_ransac");
    *status_out = 0;
    if (n > 0) {
        PLP_REQUIRE(inlier_out, "inlier_out");
        memset(inlier_out, 0, (size_t)n);
    }
    if (n == 0) return PLP_OK;                 // planar_mapping_module.cc:423-426 / :597-600
    if (n < cfg->points_per_ransac) {          // :428-436 / :602-606 (update: plane->set_invalid())
        *status_out = cfg->mode == 1 ? 2 : 0;
        return PLP_OK;
    }
    PLP_REQUIRE(pos_w && eq_inout && plane_error_inout, "null pointer");
    if (num_iter == 0) return PLP_OK;          // no iteration: best_found stays false
    PLP_REQUIRE(samples && sample_size >= 1, "samples");
    for (long long i = 0; i < (long long)num_iter * sample_size; ++i)
        PLP_REQUIRE(samples[i] >= 0 && samples[i] < n, "sample index out of range");
    PLP_CUDA_TRY(cudaSetDevice(ctx->device));
    Packer pk;
    const size_t N = (size_t)n, K = (size_t)num_iter;
    const size_t o_pos = pk.add(pos_w, N * 24), o_val = pk.add(valid, N);
    const size_t o_smp = pk.add(samples, K * (size_t)sample_size * 4);
    const size_t o_eq = pk.add(eq_inout, 32), o_pe = pk.add(plane_error_inout, 8);
    const size_t o_eqs = pk.reserve(K * 32), o_eqr = pk.reserve(K * 32), o_res = pk.reserve(K * 8), o_err = pk.reserve(K * 8);
    const size_t o_el = pk.reserve(K * 4), o_cnt = pk.reserve(K * 4), o_flag = pk.reserve(K * N), o_idx = pk.reserve(K * N * 4);
    const size_t o_inl = pk.reserve(N), o_st = pk.reserve(4);
    uint8_t *d;
    PLP_TRY(pk.upload(ctx, 0, &d));
    PlaneJob J;
    J.pos = Packer::at<double>(d, o_pos);
    J.valid = Packer::at<uint8_t>(d, o_val);
    J.samples = Packer::at<int32_t>(d, o_smp);
    J.n = n;
    J.num_iter = num_iter;
    J.sample_size = sample_size;
    J.cfg = *cfg;
    J.eq_s = Packer::at<double>(d, o_eqs);
    J.eq_r = Packer::at<double>(d, o_eqr);
    J.res = Packer::at<double>(d, o_res);
    J.err = Packer::at<double>(d, o_err);
    J.elig = Packer::at<int32_t>(d, o_el);
    J.cnt = Packer::at<int32_t>(d, o_cnt);
    J.flag = Packer::at<uint8_t>(d, o_flag);
    J.idx = Packer::at<int32_t>(d, o_idx);
    J.eq = Packer::at<double>(d, o_eq);
    J.plane_err = Packer::at<double>(d, o_pe);
    J.inlier = Packer::at<uint8_t>(d, o_inl);
    J.status = Packer::at<int32_t>(d, o_st);
    PLP_LAUNCH(ctx, plane_hypothesis_kernel, num_iter, kPlThreads, 0, J);
    PLP_CHECK_LAUNCH();
    PLP_LAUNCH(ctx, plane_select_kernel, 1, kPlThreads, 0, J);
    PLP_CHECK_LAUNCH();
    PLP_CUDA_TRY(cudaMemcpyAsync(eq_inout, d + o_eq, 32, cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaMemcpyAsync(plane_error_inout, d + o_pe, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaMemcpyAsync(inlier_out, d + o_inl, N, cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaMemcpyAsync(status_out, d + o_st, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PLP_OK;
}

}  // extern "C"
