// ba_chol.cu -- reduced camera system in HBM: blocked FP64 Cholesky for bundle adjustments with more than kBaMaxFree
// non-fixed keyframes (optimize::global_bundle_adjuster::optimize after a loop closure,
// optimize/global_bundle_adjuster.cc:64-253, module/loop_bundle_adjuster.cc:81-82; large local windows).
//
// The landmark-marginalised system S dp = g (S = Hpp - sum_l Hpl Dinv Hpl^T, 6N x 6N) does not fit shared memory beyond 32
// keyframes, so it is kept as a dense lower-triangular matrix in HBM (11.5 MB at N = 200: L2-resident) and factored by a
// right-looking blocked Cholesky, 32 columns per step, three launches per block column:
//   ba_chol_diag_kernel    1 CTA     32 x 32 diagonal block in shared memory (g2o's LinearSolver: any exact SPD solve)
//   ba_chol_panel_kernel   rows/128  rows below: X L_kk^T = A_panel, one thread per row against the shared diagonal block
//   ba_chol_trail_kernel   tiles     C -= P_i P_j^T on 32 x 32 tiles with the FP64 TENSOR CORES (mma.sync m8n8k4 f64: a warp
//                                    owns an 8 x 8 tile, 8 DMMA per 32-wide panel) -- the one GEMM-shaped step of this path
// The right-hand side g rides along as row n of the matrix, so the forward substitution falls out of the factorisation;
// ba_chol_finish_kernel (1 CTA) does the back substitution L^T dp = z, the trial poses and the LM bookkeeping terms.
// Everything is enqueued on the context stream (inside the CUDA graph of one LM try); kernels return at once while the LM
// state machine is not in a solve (lambda initialisation try, finished optimize()).
#include "ba_kernels.cuh"

namespace plp {

namespace {

constexpr int kNb = 32;  // block size (columns per step)

// ---- build the dense matrix from the packed block-upper-triangular system, or initialise lambda --------------------
__global__ void ba_chol_prepare_kernel(BaDev B) {
    BaState &ST = *B.state;
    if (ST.phase == kBaDone) return;
    const int N = B.n_free, n = 6 * N, nS = B.n_pairs * 36;
    const double *packed = B.packed;
    if (ST.phase == kBaNeedInit) {  // computeLambdaInit: tau * max diag over every active vertex
        if (blockIdx.x != 0) return;  // one CTA does the (tiny) lambda initialisation
        __shared__ double s_max[256];
        double md = 0;
        for (int h = threadIdx.x; h < N; h += blockDim.x) {
            const size_t p = (size_t)h * N - (size_t)h * (h - 1) / 2;  // diagonal block (h, h)
            for (int a = 0; a < 6; ++a) md = fmax(md, fabs(packed[p * 36 + a * 7]));
        }
        s_max[threadIdx.x] = md;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < blockDim.x; ++i) md = fmax(md, s_max[i]);
            for (int w = 0; w < B.world; ++w) md = fmax(md, packed[nS + 2 * n + 1 + w]);
            ST.lambda = 1e-5 * md;
            ST.ni = 2;
            ST.phase = kBaRunning;
            ST.iter_start = 1;
            ST.have_trial = 0;
            ST.solve_active = 0;
        }
        return;
    }
    // dense lower triangle (row-major, leading dimension n) + right-hand side as row n
    const double lambda = ST.lambda;
    const size_t total = (size_t)B.n_pairs * 36;
    for (size_t e = threadIdx.x + (size_t)blockIdx.x * blockDim.x; e < total; e += (size_t)blockDim.x * gridDim.x) {
        const int p = (int)(e / 36), rc = (int)(e - (size_t)p * 36), r = rc / 6, c = rc - r * 6;
        const int bi = B.pair_bi[p], bj = B.pair_bj[p];
        const int gi = bi * 6 + r, gj = bj * 6 + c;
        double v = packed[e];
        if (bi == bj) {
            if (c > r) continue;
            if (r == c) v += lambda;
            B.dense[(size_t)gi * n + gj] = v;
        } else {
            B.dense[(size_t)gj * n + gi] = v;  // bi < bj: entry (gi, gj) of the upper part -> (gj, gi) of the lower part
        }
    }
    for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < n; i += blockDim.x * gridDim.x) B.dense[(size_t)n * n + i] = packed[nS + i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ST.solve_active = 1;
        ST.ok2 = 1;
    }
}

// ---- diagonal block --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNb * kNb) ba_chol_diag_kernel(BaDev B, int k0) {
    BaState &ST = *B.state;
    if (ST.phase == kBaDone || !ST.solve_active) return;
    const int n = 6 * B.n_free, nb = min(kNb, n - k0);
    __shared__ double L[kNb][kNb + 1];
    __shared__ int s_ok;
    const int r = threadIdx.x / kNb, c = threadIdx.x % kNb;
    if (r < nb && c <= r) L[r][c] = B.dense[(size_t)(k0 + r) * n + k0 + c];
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        if (threadIdx.x == 0) {
            const double d = L[j][j];
            if (!(d > 0.0) || !isfinite(d)) s_ok = 0;
            L[j][j] = sqrt(d);
        }
        __syncthreads();
        if (!s_ok) break;  // uniform
        if (c == j && r > j && r < nb) L[r][j] /= L[j][j];
        __syncthreads();
        if (r > j && r < nb && c > j && c <= r) L[r][c] -= L[r][j] * L[c][j];
        __syncthreads();
    }
    if (!s_ok) {
        if (threadIdx.x == 0) ST.ok2 = 0;  // not positive definite: the step is rejected (temp_chi = max), like g2o
        return;
    }
    if (r < nb && c <= r) B.dense[(size_t)(k0 + r) * n + k0 + c] = L[r][c];
}

// ---- panel: rows below the diagonal block (and the right-hand-side row n) ------------------------------------------------
__global__ void __launch_bounds__(128) ba_chol_panel_kernel(BaDev B, int k0) {
    const BaState &ST = *B.state;
    if (ST.phase == kBaDone || !ST.solve_active || !ST.ok2) return;
    const int n = 6 * B.n_free, nb = min(kNb, n - k0);
    __shared__ double L[kNb][kNb + 1];
    for (int i = threadIdx.x; i < nb * nb; i += blockDim.x) {
        const int r = i / nb, c = i - r * nb;
        if (c <= r) L[r][c] = B.dense[(size_t)(k0 + r) * n + k0 + c];
    }
    __syncthreads();
    const int row = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (row > n) return;
    double *a = B.dense + (size_t)row * n + k0;
    double x[kNb];
#pragma unroll 4
    for (int c = 0; c < nb; ++c) {
        double v = a[c];
        for (int m = 0; m < c; ++m) v -= x[m] * L[c][m];
        x[c] = v / L[c][c];
    }
    for (int c = 0; c < nb; ++c) a[c] = x[c];
}

// ---- trailing update with FP64 tensor cores ----------------------------------------------------------------------------
// D (8x8) = A (8x4) * B (4x8) + C; lane l holds a = A[l / 4][l % 4], b = B[l % 4][l / 4], c0 / c1 = C[l / 4][2 (l % 4) + {0, 1}]
__device__ __forceinline__ void dmma8x8x4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(512) ba_chol_trail_kernel(BaDev B, int k0) {
    const BaState &ST = *B.state;
    if (ST.phase == kBaDone || !ST.solve_active || !ST.ok2) return;
    const int n = 6 * B.n_free, nb = min(kNb, n - k0);
    const int t0 = k0 + nb;               // first trailing row / column
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj > ti) return;                  // lower block triangle
    const int i0 = t0 + ti * kNb, j0 = t0 + tj * kNb;
    if (i0 > n || j0 >= n) return;        // row n (the right-hand side) is part of the rows, not of the columns
    __shared__ double Pi[kNb][kNb + 1], Pj[kNb][kNb + 1];   // panel rows of the tile's rows / columns; zero-padded
    for (int e = threadIdx.x; e < kNb * kNb; e += blockDim.x) {
        const int r = e / kNb, c = e - r * kNb;
        Pi[r][c] = (i0 + r <= n && c < nb) ? B.dense[(size_t)(i0 + r) * n + k0 + c] : 0.0;
        Pj[r][c] = (j0 + r < n && c < nb) ? B.dense[(size_t)(j0 + r) * n + k0 + c] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = (warp >> 2) * 8, wc = (warp & 3) * 8;    // 16 warps: a 4 x 4 grid of 8 x 8 tiles
    const int ar = lane >> 2, ak = lane & 3;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int k = 0; k < kNb; k += 4) dmma8x8x4(c0, c1, Pi[wr + ar][k + ak], Pj[wc + ar][k + ak]);  // B[k][ncol] = Pj[ncol][k]
    const int gi = i0 + wr + ar, gj = j0 + wc + 2 * ak;
    if (gi <= n) {
        if (gj < n && gj <= gi) B.dense[(size_t)gi * n + gj] -= c0;
        if (gj + 1 < n && gj + 1 <= gi) B.dense[(size_t)gi * n + gj + 1] -= c1;
    }
}

// ---- back substitution, trial poses, LM bookkeeping (the tail of ba_solve_kernel for the dense-in-HBM system) -----------
__global__ void __launch_bounds__(1024) ba_chol_finish_kernel(BaDev B) {
    BaState &ST = *B.state;
    if (ST.phase == kBaDone || !ST.solve_active) return;
    extern __shared__ __align__(16) double fs[];
    const int tid = threadIdx.x, n = 6 * B.n_free, nS = B.n_pairs * 36;
    double *z = fs, *x = fs + n;   // z = L^-1 g (row n of the factored matrix), then the solution
    const int ok = ST.ok2;
    for (int i = tid; i < n; i += blockDim.x) {
        z[i] = ok ? B.dense[(size_t)n * n + i] : 0.0;
        x[i] = 0.0;
    }
    __syncthreads();
    if (ok) {
        // L^T x = z, block by block from the bottom; inside a block one warp resolves the 32 unknowns in sequence, then all
        // threads subtract the block's contribution from the rows above (row-major L: coalesced over the columns)
        __shared__ double Lb[kNb][kNb + 1];
        for (int k0 = ((n - 1) / kNb) * kNb; k0 >= 0; k0 -= kNb) {
            const int nb = min(kNb, n - k0);
            for (int e = tid; e < nb * nb; e += blockDim.x) {
                const int r = e / nb, c = e - r * nb;
                if (c <= r) Lb[r][c] = B.dense[(size_t)(k0 + r) * n + k0 + c];
            }
            __syncthreads();
            if (tid < 32) {
                for (int j = nb - 1; j >= 0; --j) {
                    double xj = 0;
                    if (tid == 0) {
                        xj = z[k0 + j] / Lb[j][j];
                        x[k0 + j] = xj;
                    }
                    xj = __shfl_sync(0xffffffffu, xj, 0);
                    if (tid < j) z[k0 + tid] -= Lb[j][tid] * xj;
                    __syncwarp();
                }
            }
            __syncthreads();
            for (int i = tid; i < k0; i += blockDim.x) {
                double s = 0;
                for (int j = 0; j < nb; ++j) s += B.dense[(size_t)(k0 + j) * n + i] * x[k0 + j];
                z[i] -= s;
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += blockDim.x) B.dp[i] = x[i];
    const int cur = ST.cur;
    for (int k = tid; k < B.n_kf; k += blockDim.x) {
        const int h = B.kf_hidx[k];
        se3::Pose P = B.poses[cur][k];
        if (h >= 0 && ok) P = se3::oplus(P, x + 6 * h);
        B.poses[cur ^ 1][k] = P;
    }
    __syncthreads();
    // scale term dp^T (lambda dp + bp): block sum
    __shared__ double s_part[32];
    double sc = 0;
    const double lambda = ST.lambda;
    for (int i = tid; i < n; i += blockDim.x) sc += x[i] * (lambda * x[i] + B.packed[nS + n + i]);
    for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
    if ((tid & 31) == 0) s_part[tid >> 5] = sc;
    __syncthreads();
    if (tid == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
        if (ST.iter_start) {  // currentChi = activeRobustChi2() at the start of an iteration
            ST.current_chi = B.packed[nS + 2 * n];
            ST.qmax = 0;
            ST.iter_start = 0;
        }
        ST.scale_pose = t;
        ST.have_trial = 1;
    }
}

}  // namespace

size_t ba_dense_bytes(int n_free) {
    const size_t n = 6 * (size_t)n_free;
    return (n + 1) * n * sizeof(double);
}

plp_status ba_launch_solve_large(plp_ctx *ctx, const BaDev &B) {
    const int n = 6 * B.n_free;
    PLP_LAUNCH(ctx, ba_chol_prepare_kernel, B.phase_init_grid, 256, 0, B);
    for (int k0 = 0; k0 < n; k0 += kNb) {
        const int nb = std::min(kNb, n - k0);
        PLP_LAUNCH(ctx, ba_chol_diag_kernel, 1, kNb * kNb, 0, B, k0);
        const int rows = n + 1 - (k0 + nb);
        if (rows > 0) {
            PLP_LAUNCH(ctx, ba_chol_panel_kernel, div_up(rows, 128), 128, 0, B, k0);
            const int tiles = div_up(rows, kNb);
            PLP_LAUNCH(ctx, ba_chol_trail_kernel, dim3(tiles, tiles), 512, 0, B, k0);
        }
    }
    PLP_LAUNCH(ctx, ba_chol_finish_kernel, 1, 1024, (size_t)2 * n * sizeof(double), B);
    PLP_CHECK_LAUNCH();
    return PLP_OK;
}

}  // namespace plp
