// ba_nccl.cu -- the only collective of the framework: one sum-all-reduce (f64) of the packed reduced camera
// system [S | g | bp | chi2 | max-diag slots] per LM try, plus the 2-double {chi2_trial, scale} reduction that the
// rho test needs -- through a one-shot kernel over NVLink peer memory when the ranks share a node, else ncclAllReduce (see DESIGN.md section 6 for why the exact g2o accept/reject rule needs
// that second, 16-byte reduction).  Landmarks are sharded over ranks, the <= 32 free poses are replicated and every
// rank solves the identical reduced system redundantly (deterministic, no broadcast).
#include <nccl.h>

#include "ba_kernels.cuh"

using namespace plp;

// ---------------------------------------------------------------------------------------------------------------------
// One-shot all-reduce over NVLink peer memory (single node, <= 8 ranks, vectors of at most kPeerCap doubles).
//
// The vectors of this path are small (62 KB at config 4) and a try cannot go on without the sum, so the cost of a collective is
// its LATENCY: ncclAllReduce takes ~35-60 us at 8 ranks for this size, more than every other kernel of the try.  Here each
// rank owns a mailbox in its HBM that the peers map through CUDA IPC.  A call: every CTA copies its chunk of the local
// vector into the rank's own mailbox, fences, stores the call's sequence number into the chunk's flag word on every peer
// (NVLink stores), waits until all peers' numbers have arrived in its own flag words (local polling), then reads the chunk
// from all mailboxes over NVLink and adds in rank order -- every rank forms the sums in the same order, so the replicated
// LM state stays bit-identical.  Data is double-buffered by the parity of the sequence number: a mailbox half is rewritten
// two calls later, which a rank can only reach after every peer has signalled the call in between, i.e. after every peer
// has finished reading.  The sequence number lives in device memory (the last CTA of a call advances it), so the call can be
// captured in the CUDA graph of an LM try.  A peer that never shows up ends the wait after 2 s and raises the error word.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kPeerMaxWorld = 8;
constexpr int kPeerCap = 1 << 16;   // doubles per mailbox half (512 KB)
constexpr int kPeerChunks = 16;     // CTAs per call = independent (chunk, flag) pipelines
constexpr int kPeerThreads = 512;

struct PeerArgs {
    double *mail[kPeerMaxWorld];               // [2][kPeerCap] of every rank (own + IPC-mapped)
    unsigned long long *flags[kPeerMaxWorld];  // [kPeerChunks][kPeerMaxWorld] of every rank
    unsigned long long *seq;                   // device-resident call counter (local)
    unsigned *done;                            // CTAs finished in this call (local)
    int *err;                                  // sticky error word (local)
    int world, rank;
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(kPeerThreads) ba_peer_allreduce_kernel(PeerArgs A, double *buf, int n) {
    const unsigned long long seq = *A.seq + 1;
    const int parity = (int)(seq & 1ull);
    const int tid = threadIdx.x, c = blockIdx.x;
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int i0 = min(n, c * per), i1 = min(n, i0 + per);
    double *mine = A.mail[A.rank] + (size_t)parity * kPeerCap;
    for (int i = i0 + tid; i < i1; i += kPeerThreads) mine[i] = buf[i];
    __syncthreads();
    if (tid < A.world) {
        // the barrier made the CTA's stores visible to this thread; its system-scope fence orders them before the signal
        __threadfence_system();
        // signal: "rank A.rank has published chunk c of call seq" on every rank (own included)
        *reinterpret_cast<volatile unsigned long long *>(A.flags[tid] + c * kPeerMaxWorld + A.rank) = seq;
        // wait for every rank's signal in the local flag words
        volatile unsigned long long *f = A.flags[A.rank] + c * kPeerMaxWorld + tid;
        const unsigned long long t0 = global_timer_ns();
        while (*f < seq) {
            if (global_timer_ns() - t0 > 2000000000ull) {
                *A.err = 1;
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    for (int i = i0 + tid; i < i1; i += kPeerThreads) {
        double s = 0.0;
        for (int r = 0; r < A.world; ++r) s += __ldcv(A.mail[r] + (size_t)parity * kPeerCap + i);  // fixed order on every rank
        buf[i] = s;
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(A.done, 1u) == gridDim.x - 1) {  // every CTA has read `seq`: advance it for the next call
            *A.done = 0u;
            *A.seq = seq;
        }
    }
}

struct plp_ba_comm : public BaCollective {
    plp_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    uint64_t calls = 0;  // all-reduces issued through this communicator (bench.py: all-reduces per LM try)
    uint64_t peer_calls = 0;
    // NVLink peer path
    bool peer_ok = false;
    uint8_t *d_mail = nullptr;  // own mailbox block: data, flags, seq, done, err
    void *peer_base[kPeerMaxWorld] = {nullptr};
    PeerArgs args{};

    bool use_peer(int n) const { return peer_ok && n <= kPeerCap; }
    bool graph_safe(int n_max) const override { return use_peer(n_max); }
    void add_calls(uint64_t n) override {
        calls += n;
        peer_calls += n;
    }
    plp_status all_reduce(double *d_buf, int n) override {
        if (use_peer(n)) {
            const int grid = std::max(1, std::min(kPeerChunks, (n + kPeerThreads - 1) / kPeerThreads));
            ba_peer_allreduce_kernel<<<grid, kPeerThreads, 0, ctx->stream>>>(args, d_buf, n);
            ctx->launches++;
            // while the LM try is being captured into its CUDA graph nothing runs: the replays are counted by add_calls()
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(ctx->stream, &cs) != cudaSuccess || cs == cudaStreamCaptureStatusNone) {
                calls++;
                peer_calls++;
            }
            return PLP_OK;
        }
        const ncclResult_t r = ncclAllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, comm, ctx->stream);
        if (r != ncclSuccess) {
            set_error("ncclAllReduce failed: %s", ncclGetErrorString(r));
            return PLP_ERR_NCCL;
        }
        ctx->launches++;  // the NCCL kernel
        calls++;
        return PLP_OK;
    }
    plp_status check() override {  // after a stream synchronisation: did a peer wait time out?
        if (!peer_ok) return PLP_OK;
        int e = 0;
        PLP_CUDA_TRY(cudaMemcpy(&e, args.err, sizeof(int), cudaMemcpyDeviceToHost));
        if (e) {
            set_error("NVLink peer all-reduce: a rank did not arrive within 2 s");
            return PLP_ERR_NCCL;
        }
        return PLP_OK;
    }
};

// mailbox block layout
static constexpr size_t kMailData = (size_t)2 * kPeerCap * sizeof(double);
static constexpr size_t kMailFlags = (size_t)kPeerChunks * kPeerMaxWorld * sizeof(unsigned long long);
static constexpr size_t kMailBytes = kMailData + kMailFlags + 256;

// Exchange the IPC handles of the mailboxes through the NCCL communicator that already exists and map the peers.
// Any failure leaves the communicator on its NCCL path (peer_ok = false): the result is the same, only slower.
static void peer_setup(plp_ba_comm *c) {
    const char *ev = getenv("PLP_BA_PEER");
    if (ev && atoi(ev) == 0) return;
    if (c->world < 2 || c->world > kPeerMaxWorld) return;
    cudaStream_t st = c->ctx->stream;
    cudaIpcMemHandle_t mine, *all = nullptr;
    uint8_t *d_h = nullptr;
    bool ok = cudaMalloc((void **)&c->d_mail, kMailBytes) == cudaSuccess && cudaMemsetAsync(c->d_mail, 0, kMailBytes, st) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine, c->d_mail) == cudaSuccess &&
              cudaMalloc((void **)&d_h, sizeof(mine) * (size_t)(c->world + 1)) == cudaSuccess;
    // every rank takes part in the all-gather even if its own setup failed (a zero handle marks the failure)
    if (!ok) memset(&mine, 0, sizeof(mine));
    all = (cudaIpcMemHandle_t *)malloc(sizeof(mine) * (size_t)c->world);
    bool gathered = false;
    if (d_h && all) {
        gathered = cudaMemcpyAsync(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice, st) == cudaSuccess &&
                   ncclAllGather(d_h, d_h + sizeof(mine), sizeof(mine), ncclUint8, c->comm, st) == ncclSuccess &&
                   cudaMemcpyAsync(all, d_h + sizeof(mine), sizeof(mine) * (size_t)c->world, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                   cudaStreamSynchronize(st) == cudaSuccess;
    }
    if (gathered) {
        const cudaIpcMemHandle_t zero{};
        for (int r = 0; r < c->world && ok; ++r) {
            if (memcmp(&all[r], &zero, sizeof(zero)) == 0) ok = false;
        }
        for (int r = 0; r < c->world && ok; ++r) {
            if (r == c->rank) {
                c->peer_base[r] = c->d_mail;
            } else if (cudaIpcOpenMemHandle(&c->peer_base[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                c->peer_base[r] = nullptr;
                ok = false;
            }
        }
    } else {
        ok = false;
    }
    // the decision must be unanimous: a rank that cannot map a peer would otherwise wait for signals nobody sends
    int *d_ok = (int *)d_h;
    int h_ok = ok ? 1 : 0;
    if (d_h && cudaMemcpyAsync(d_ok, &h_ok, sizeof(int), cudaMemcpyHostToDevice, st) == cudaSuccess &&
        ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, c->comm, st) == ncclSuccess &&
        cudaMemcpyAsync(&h_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess) {
        ok = h_ok == 1;
    } else {
        ok = false;
    }
    if (d_h) cudaFree(d_h);
    free(all);
    if (!ok) {
        cudaGetLastError();
        return;
    }
    for (int r = 0; r < c->world; ++r) {
        c->args.mail[r] = reinterpret_cast<double *>(c->peer_base[r]);
        c->args.flags[r] = reinterpret_cast<unsigned long long *>(reinterpret_cast<uint8_t *>(c->peer_base[r]) + kMailData);
    }
    uint8_t *tail = c->d_mail + kMailData + kMailFlags;
    c->args.seq = reinterpret_cast<unsigned long long *>(tail);
    c->args.done = reinterpret_cast<unsigned *>(tail + 64);
    c->args.err = reinterpret_cast<int *>(tail + 128);
    c->args.world = c->world;
    c->args.rank = c->rank;
    c->peer_ok = true;
}

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
    peer_setup(c);
    *out = c;
    return PLP_OK;
}

int plp_ba_comm_peer_active(const plp_ba_comm *c) { return (c && c->peer_ok) ? 1 : 0; }
uint64_t plp_ba_comm_peer_count(const plp_ba_comm *c) { return c ? c->peer_calls : 0; }

uint64_t plp_ba_comm_allreduce_count(const plp_ba_comm *c) { return c ? c->calls : 0; }

void plp_ba_comm_destroy(plp_ba_comm *c) {
    if (!c) return;
    if (c->ctx) {
        cudaSetDevice(c->ctx->device);
        cudaStreamSynchronize(c->ctx->stream);
    }
    for (int r = 0; r < c->world && r < kPeerMaxWorld; ++r)
        if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
    if (c->comm) ncclCommDestroy(c->comm);  // (a barrier: no peer reads this rank's mailbox any more)
    if (c->d_mail) cudaFree(c->d_mail);
    delete c;
}

}  // extern "C"
