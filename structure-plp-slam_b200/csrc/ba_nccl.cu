// ba_nccl.cu -- the only collective of the framework: one ncclAllReduce(sum, f64) of the packed reduced camera
// system [S | g | bp | chi2 | max-diag slots] per LM try, plus the 2-double {chi2_trial, scale} reduction that the
// rho test needs (see DESIGN.md section 6 for why the exact g2o accept/reject rule needs
// that second, 16-byte reduction).  Landmarks are sharded over ranks, the <= 32 free poses are replicated and every
// rank solves the identical reduced system redundantly (deterministic, no broadcast).
#include <nccl.h>

#include "ba_kernels.cuh"

using namespace plp;

struct plp_ba_comm : public BaCollective {
    plp_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    uint64_t calls = 0;  // ncclAllReduce calls issued through this communicator (bench.py: all-reduces per LM try)
    plp_status all_reduce(double *d_buf, int n) override {
        const ncclResult_t r = ncclAllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, comm, ctx->stream);
        if (r != ncclSuccess) {
            set_error("ncclAllReduce failed: %s", ncclGetErrorString(r));
            return PLP_ERR_NCCL;
        }
        ctx->launches++;  // the NCCL kernel
        calls++;
        return PLP_OK;
    }
};

namespace plp {
BaCollective *ba_comm_collective(plp_ba_comm *c) { return c; }
int ba_comm_rank(plp_ba_comm *c) { return c->rank; }
int ba_comm_world(plp_ba_comm *c) { return c->world; }
}  // namespace plp

extern "C" {

plp_status plp_ba_comm_unique_id(uint8_t id_out[128]) {
    PLP_REQUIRE(id_out != nullptr, "id_out");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    const ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess) {
        set_error("ncclGetUniqueId failed: %s", ncclGetErrorString(r));
        return PLP_ERR_NCCL;
    }
    memcpy(id_out, &id, 128);
    return PLP_OK;
}

plp_status plp_ba_comm_init(plp_ctx *ctx, const uint8_t id[128], int world, int rank, plp_ba_comm **out) {
    PLP_REQUIRE(ctx && id && out && world >= 1 && rank >= 0 && rank < world, "args");
    *out = nullptr;
    PLP_CUDA_TRY(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    plp_ba_comm *c = new plp_ba_comm();
    c->ctx = ctx;
    c->world = world;
    c->rank = rank;
    const ncclResult_t r = ncclCommInitRank(&c->comm, world, uid, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", ncclGetErrorString(r));
        delete c;
        return PLP_ERR_NCCL;
    }
    *out = c;
    return PLP_OK;
}

uint64_t plp_ba_comm_allreduce_count(const plp_ba_comm *c) { return c ? c->calls : 0; }

void plp_ba_comm_destroy(plp_ba_comm *c) {
    if (!c) return;
    if (c->comm) ncclCommDestroy(c->comm);
    delete c;
}

}  // extern "C"
