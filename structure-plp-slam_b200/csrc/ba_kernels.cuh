// ba_kernels.cuh -- device-side problem descriptor and launch helpers of the local bundle adjuster
// (local_ba.cu = kernels, ba_host.cu = C ABI + LM driver, ba_nccl.cu = multi-GPU collective).
#pragma once
#include "common.cuh"
#include "se3.cuh"
#include "ba_types.cuh"

namespace plp {

// multi-GPU hook: sum-all-reduce of `n` doubles in place on the context stream (ba_nccl.cu)
struct BaCollective {
    virtual plp_status all_reduce(double *d_buf, int n) = 0;
    virtual bool graph_safe(int n_max) const { return false; }  // may all_reduce(<= n_max doubles) be captured in a CUDA graph?
    virtual void add_calls(uint64_t n) {}                       // a captured graph with n all-reduces was replayed
    virtual plp_status check() { return PLP_OK; }               // after a stream synchronisation
    virtual ~BaCollective() {}
};

size_t ba_linearize_smem(int n_free, int n_pairs, int pool_cap);
int ba_pool_capacity(int n_free, int n_pairs, int max_free_degree);
size_t ba_solve_smem(int n_free);
plp_status ba_prepare_kernels(int n_free, int n_pairs, int pool_cap);
plp_status ba_launch_try(plp_ctx *ctx, const BaDev &B, BaCollective *coll);
size_t ba_dense_bytes(int n_free);
plp_status ba_launch_solve_large(plp_ctx *ctx, const BaDev &B);
plp_status ba_launch_decide(plp_ctx *ctx, const BaDev &B);
plp_status ba_launch_set_state(plp_ctx *ctx, const BaDev &B, int max_it, int robust, int reset_cur);
plp_status ba_launch_classify(plp_ctx *ctx, const BaDev &B, int set_levels);
plp_status ba_launch_init_poses(plp_ctx *ctx, const BaDev &B, const double *d_T_in);
plp_status ba_launch_export(plp_ctx *ctx, const BaDev &B, double *d_T_out, double *d_pts_out, double *d_lines_out);

}  // namespace plp
