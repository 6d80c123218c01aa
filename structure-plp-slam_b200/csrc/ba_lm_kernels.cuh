// ba_lm_kernels.cuh -- device code of the Schur-complement Levenberg-Marquardt bundle adjuster (see local_ba.cu for the
// work decomposition and the reference it replaces).  Free of host-side CUDA runtime dependencies so that tests/cta_emu can
// compile the same text for the host and run whole LM tries against the oracle on the CPU.
#pragma once
#include <math.h>
#include <stdint.h>

#include "ba_types.cuh"
#include "se3.cuh"

namespace plp {
namespace balm {

using se3::Pose;

constexpr int kBaThreads = 512;
constexpr int kBaWarps = kBaThreads / 32;
constexpr int kPoolMax = kBaWarps * kBaMaxFree;  // upper bound of the per-batch pool of (free keyframe, landmark) blocks
constexpr double kDelta = 1e-9;  // g2o numeric Jacobian step

// ---------------------------------------------------------------------------------------------------------
// Line3D (optimize/g2o/line3d.h:57-207): Pluecker (w, d) <-> orthonormal (U in SO3, W in SO2)
// ---------------------------------------------------------------------------------------------------------
__device__ void line_oplus(const double *L, const double *v, double *out) {
    // toOrthonormal (line3d.h:137-157)
    const double mx = sqrt(L[3] * L[3] + L[4] * L[4] + L[5] * L[5]);  // |d|
    const double my = sqrt(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);  // |w|
    const double wn = 1.0 / sqrt(mx * mx + my * my);
    double W[4] = {my * wn, -mx * wn, mx * wn, my * wn};
    const double mn = 1.0 / my, dn = 1.0 / mx;
    const double cx = L[1] * L[5] - L[2] * L[4], cy = L[2] * L[3] - L[0] * L[5], cz = L[0] * L[4] - L[1] * L[3];
    const double cn = 1.0 / sqrt(cx * cx + cy * cy + cz * cz);
    double U[9] = {L[0] * mn, L[3] * dn, cx * cn, L[1] * mn, L[4] * dn, cy * cn, L[2] * mn, L[5] * dn, cz * cn};
    // update (line3d.h:171-186)
    const double c = cos(v[3]), s = sin(v[3]);
    double qw = sqrt(1 - (v[0] * v[0] + v[1] * v[1] + v[2] * v[2])), qx = v[0], qy = v[1], qz = v[2];
    const double qn = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
    qw /= qn;
    qx /= qn;
    qy /= qn;
    qz /= qn;
    double Ru[9];
    se3::quat_to_R(qw, qx, qy, qz, Ru);
    double U2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) U2[i * 3 + j] = U[i * 3] * Ru[j] + U[i * 3 + 1] * Ru[3 + j] + U[i * 3 + 2] * Ru[6 + j];
    const double W0 = W[0] * c + W[1] * s, W2 = W[2] * c + W[3] * s;
    // fromOrthonormal (line3d.h:116-134) + normalize (twice, as in oplus)
    double o[6] = {U2[0] * W0, U2[3] * W0, U2[6] * W0, U2[1] * W2, U2[4] * W2, U2[7] * W2};
    for (int rep = 0; rep < 2; ++rep) {
        const double n = 1.0 / sqrt(o[3] * o[3] + o[4] * o[4] + o[5] * o[5]);
        for (int k = 0; k < 6; ++k) o[k] *= n;
    }
    for (int k = 0; k < 6; ++k) out[k] = o[k];
}

// reproj_edge_line3d::depth_is_positive_via_endpoints_trimming (reproj_edge_line3d_orthonormal.h:97-177)
__device__ bool line_depth_positive(const se3::Cam &c, const Pose &P, const double *L, const float *obs) {
    const double *R = P.R, *t = P.t;
    double Rn[3], Rd[3];
    for (int r = 0; r < 3; ++r) {
        Rn[r] = R[r * 3] * L[0] + R[r * 3 + 1] * L[1] + R[r * 3 + 2] * L[2];
        Rd[r] = R[r * 3] * L[3] + R[r * 3 + 1] * L[4] + R[r * 3 + 2] * L[5];
    }
    const double lc0 = Rn[0] + (t[1] * Rd[2] - t[2] * Rd[1]), lc1 = Rn[1] + (t[2] * Rd[0] - t[0] * Rd[2]);
    const double lc2 = Rn[2] + (t[0] * Rd[1] - t[1] * Rd[0]);
    const double l1 = c.fy * lc0, l2 = c.fx * lc1, l3 = -c.fy * c.cx * lc0 - c.fx * c.cy * lc1 + c.fx * c.fy * lc2;
    const double sp0 = obs[0], sp1 = obs[1], ep0 = obs[2], ep1 = obs[3];
    const double x_sp = -(sp1 - (l2 / l1) * sp0 + (l3 / l2)) * ((l1 * l2) / (l1 * l1 + l2 * l2));
    const double y_sp = -(l1 / l2) * x_sp - (l3 / l2);
    const double x_ep = -(ep1 - (l2 / l1) * ep0 + (l3 / l2)) * ((l1 * l2) / (l1 * l1 + l2 * l2));
    const double y_ep = -(l1 / l2) * x_ep - (l3 / l2);
    const double y_0sp = sp1 - (l2 / l1) * sp0, y_0ep = ep1 - (l2 / l1) * ep0;
    double Pm[12];
    const double K[9] = {c.fx, 0, c.cx, 0, c.fy, c.cy, 0, 0, 1};
    for (int r = 0; r < 3; ++r)
        for (int col = 0; col < 4; ++col) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += K[r * 3 + k] * (col < 3 ? R[k * 3 + col] : t[k]);
            Pm[r * 4 + col] = s;
        }
    double depth[2];
    for (int which = 0; which < 2; ++which) {
        const double xc = which ? x_ep : x_sp, yc = which ? y_ep : y_sp, y0 = which ? y_0ep : y_0sp;
        // line through (xc, yc, 1) and (0, y0, 1)
        const double a0 = yc * 1.0 - 1.0 * y0, a1 = 1.0 * 0.0 - xc * 1.0, a2 = xc * y0 - yc * 0.0;
        double pl[4];
        for (int col = 0; col < 4; ++col) pl[col] = Pm[col] * a0 + Pm[4 + col] * a1 + Pm[8 + col] * a2;
        // [m]x pl.head<3> + d pl[3] ; -d . pl.head<3>
        const double X0 = (-L[2] * pl[1] + L[1] * pl[2]) + L[3] * pl[3];
        const double X1 = (L[2] * pl[0] - L[0] * pl[2]) + L[4] * pl[3];
        const double X2 = (-L[1] * pl[0] + L[0] * pl[1]) + L[5] * pl[3];
        const double X3 = -(L[3] * pl[0] + L[4] * pl[1] + L[5] * pl[2]);
        depth[which] = R[6] * (X0 / X3) + R[7] * (X1 / X3) + R[8] * (X2 / X3) + t[2] * 1.0;
    }
    return 0 < depth[0] && 0 < depth[1];
}

// inverse of a small symmetric matrix (D = 3 or 4) by Gauss-Jordan with partial pivoting
__device__ bool inv_small(const double *A, int D, double *Ai) {
    double M[16], I[16];
    for (int i = 0; i < D * D; ++i) {
        M[i] = A[i];
        I[i] = 0;
    }
    for (int i = 0; i < D; ++i) I[i * D + i] = 1;
    for (int c = 0; c < D; ++c) {
        int piv = c;
        for (int r = c + 1; r < D; ++r)
            if (fabs(M[r * D + c]) > fabs(M[piv * D + c])) piv = r;
        if (M[piv * D + c] == 0.0) return false;
        if (piv != c)
            for (int k = 0; k < D; ++k) {
                double tmp = M[c * D + k];
                M[c * D + k] = M[piv * D + k];
                M[piv * D + k] = tmp;
                tmp = I[c * D + k];
                I[c * D + k] = I[piv * D + k];
                I[piv * D + k] = tmp;
            }
        const double d = 1.0 / M[c * D + c];
        for (int k = 0; k < D; ++k) {
            M[c * D + k] *= d;
            I[c * D + k] *= d;
        }
        for (int r = 0; r < D; ++r) {
            if (r == c) continue;
            const double f = M[r * D + c];
            for (int k = 0; k < D; ++k) {
                M[r * D + k] -= f * M[c * D + k];
                I[r * D + k] -= f * I[c * D + k];
            }
        }
    }
    for (int i = 0; i < D * D; ++i) Ai[i] = I[i];
    return true;
}

__device__ __noinline__ bool inv_small_cold(const double *A, int D, double *Ai) { return inv_small(A, D, Ai); }

// ---------------------------------------------------------------------------------------------------------
// edge evaluation shared by linearize / update / classify
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double eval_pt(const se3::Cam &cam, const Pose &P, const double *X, const float *obs, double info,
                                          double *e, double *pc) {
    const bool stereo = !(obs[2] < 0);
    se3::map_point(P.R, P.t, X, pc);
    const double o[3] = {(double)obs[0], (double)obs[1], (double)obs[2]};
    se3::point_error(cam, pc, o, stereo, e);
    return e[0] * (info * e[0]) + e[1] * (info * e[1]) + (stereo ? e[2] * (info * e[2]) : 0.0);
}
__device__ __forceinline__ double eval_ln(const se3::Cam &cam, const Pose &P, const double *L, const float *obs, double info,
                                          double *e) {
    const double o[4] = {(double)obs[0], (double)obs[1], (double)obs[2], (double)obs[3]};
    se3::line_error(cam, P.R, P.t, L, o, e);
    return e[0] * (info * e[0]) + e[1] * (info * e[1]);
}
__device__ __forceinline__ double eval_plane(const double *X, const double *fn) {
    return (X[0] * fn[0] + X[1] * fn[1] + X[2] * fn[2] + fn[3]) / sqrt(fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2]);
}

struct BaPoolEntry {
    double W[24];   // Hpl block, 6 x D row-major
    double Y[24];   // W * Dinv
    double A[21];   // Jp^T w Jp, upper triangle row-wise
    double bpe[6];  // -Jp^T w e
    double gpe[6];  // bpe - W * (Dinv bl)
    int h, pad;     // free-keyframe index of the edge (large path: the accumulation walks the pool, not the slot table)
};

struct BaLineScratch {  // one line edge at a time: its numeric Jacobians, residual and weight, written by the evaluating lanes
    double jp[12], jl[8], r[2], w, pad;
};

struct BaSmem {  // the pool (B.pool_cap entries) and the partial system follow in dynamic shared memory
    unsigned kfmask[kBaMaxFree];  // per free keyframe: which landmarks of the batch observe it (bit = landmark in batch)
    double pert_line[kBaWarps][8][6];  // per warp: the 8 perturbed lines of its landmark
    BaLineScratch scr[kBaWarps];
    double hll[kBaWarps][16], bl[kBaWarps][4], dinv[kBaWarps][16], dl[kBaWarps][4];
    short slot[kBaWarps][kBaMaxFree];  // landmark-in-batch x free keyframe -> pool index (-1: none)
    int warp_cnt[kBaWarps];
    double red[kBaWarps][2];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// inverse of the symmetric positive definite D x D landmark block (Hll + lambda I) by an unrolled Cholesky factorisation:
// everything stays in registers.  Returns false when a pivot is not positive (the caller falls back to inv_small).
template <int D>
__device__ __forceinline__ bool inv_spd(const double *H /*D x D, full*/, double *Hi) {
    double l[D][D], li[D][D];  // L (lower) and L^-1 (lower)
    bool ok = true;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        double d = H[c * D + c];
#pragma unroll
        for (int m = 0; m < c; ++m) d -= l[c][m] * l[c][m];
        ok = ok && (d > 0.0) && isfinite(d);
        const double inv = 1.0 / sqrt(d);
        l[c][c] = inv;  // the diagonal keeps 1 / l_cc
#pragma unroll
        for (int r = c + 1; r < D; ++r) {
            double v = H[r * D + c];
#pragma unroll
            for (int m = 0; m < c; ++m) v -= l[r][m] * l[c][m];
            l[r][c] = v * inv;
        }
    }
    // L^-1 by forward substitution on the identity
#pragma unroll
    for (int c = 0; c < D; ++c) {
        li[c][c] = l[c][c];
#pragma unroll
        for (int r = c + 1; r < D; ++r) {
            double v = 0.0;
#pragma unroll
            for (int m = c; m < r; ++m) v -= l[r][m] * li[m][c];
            li[r][c] = v * l[r][r];
        }
    }
    // H^-1 = L^-T L^-1
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int c = a; c < D; ++c) {
            double s = 0.0;
#pragma unroll
            for (int m = c; m < D; ++m) s += li[m][a] * li[m][c];
            Hi[a * D + c] = s;
            Hi[c * D + a] = s;
        }
    return ok;
}

// lane 0 of a landmark's warp: damped block inverse, Dinv bl, the per-landmark outputs the back-substitution reads
template <int D>
__device__ __forceinline__ void finish_landmark(const BaDev &B, BaSmem &S, int warp, bool is_line, int li, const double *hll,
                                                const double *bl, bool any_active, bool init_mode, double lambda,
                                                double &maxdiag) {
    double H[D * D];
    int q = 0;
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int c = a; c < D; ++c) {
            H[a * D + c] = hll[q];
            H[c * D + a] = hll[q];
            ++q;
        }
    if (any_active) {
#pragma unroll
        for (int a = 0; a < D; ++a) maxdiag = fmax(maxdiag, fabs(H[a * D + a]));
    }
#pragma unroll
    for (int a = 0; a < D; ++a) H[a * D + a] += lambda;
    double Di[D * D];
    bool ok = false;
    if (any_active && !init_mode) {
        ok = inv_spd<D>(H, Di);
        if (!ok) {  // numerically indefinite block: pivoted Gauss-Jordan (cold; its operands live in local memory)
            double Hc[D * D], Dc[D * D];
#pragma unroll
            for (int i = 0; i < D * D; ++i) Hc[i] = H[i];
            ok = inv_small_cold(Hc, D, Dc);
#pragma unroll
            for (int i = 0; i < D * D; ++i) Di[i] = Dc[i];
        }
    }
    if (!ok) {
#pragma unroll
        for (int i = 0; i < D * D; ++i) Di[i] = 0.0;
    }
    double *Dg = (is_line ? B.ln_Dinv + 16 * (size_t)li : B.pt_Dinv + 16 * (size_t)li);
    double *bg = (is_line ? B.ln_bl + 4 * (size_t)li : B.pt_bl + 4 * (size_t)li);
#pragma unroll
    for (int i = 0; i < D * D; ++i) {
        S.dinv[warp][i] = Di[i];
        Dg[i] = Di[i];
    }
#pragma unroll
    for (int a = 0; a < D; ++a) {
        double s = 0;
#pragma unroll
        for (int c = 0; c < D; ++c) s += Di[a * D + c] * bl[c];
        S.dl[warp][a] = s;
        S.bl[warp][a] = bl[a];
        bg[a] = bl[a];
    }
    (is_line ? B.ln_active : B.pt_active)[li] = (any_active && (ok || init_mode)) ? 1 : 0;
}

// Y = W Dinv and gpe = bpe - W (Dinv bl) of a landmark's pool entries, one output per lane and round
template <int D>
__device__ __forceinline__ void pool_products(BaSmem &S, BaPoolEntry *pool, int warp, int lane, int pb, int pn) {
    constexpr int kPer = 6 * D + 6;
    const double *Di = S.dinv[warp], *dl = S.dl[warp];
    for (int idx = lane; idx < pn * kPer; idx += 32) {
        const int i = idx / kPer, o = idx - i * kPer;
        BaPoolEntry &pe = pool[pb + i];
        if (o < 6 * D) {
            const int a = o / D, c = o - a * D;
            double s = 0;
#pragma unroll
            for (int k = 0; k < D; ++k) s += pe.W[a * D + k] * Di[k * D + c];
            pe.Y[o] = s;
        } else {
            const int a = o - 6 * D;
            double gs = pe.bpe[a];
#pragma unroll
            for (int c = 0; c < D; ++c) gs -= pe.W[a * D + c] * dl[c];
            pe.gpe[a] = gs;
        }
    }
}

// =========================================================================================================
// ba_linearize_kernel
// =========================================================================================================
// kLarge = false: <= kBaMaxFree non-fixed keyframes, the CTA's share of the reduced camera system lives in shared memory
// (no atomics, fixed summation order).  kLarge = true (global BA / large local windows): the reduced system is the dense
// block-upper-triangular `packed` vector in HBM (L2-resident: 5.8 MB for 200 keyframes) and every landmark adds its
// -Y_i W_j^T / A blocks with FP64 atomics; nothing in the kernel is sized by the number of keyframes any more.
//
// Phase 1, one warp per landmark.  Point landmarks (analytic Jacobians): lane = edge.  Line landmarks (numeric Jacobians:
// 21 evaluations of the error function per edge -- the estimate, 6 x 2 perturbed poses, 4 x 2 perturbed lines): one edge
// at a time, lane = evaluation, the central differences by one shuffle, then lane = output entry of Hll / bl / W / A / b.
template <bool kLarge>
__global__ void __launch_bounds__(kBaThreads, 1) ba_linearize_kernel(BaDev B) {
#ifdef PLP_CTA_EMU
    uint8_t *ba_smem_raw = emu_dynamic_smem;
#else
    extern __shared__ __align__(16) uint8_t ba_smem_raw[];
#endif
    const BaState &ST = *B.state;
    if (ST.phase == kBaDone) return;
    BaSmem &S = *reinterpret_cast<BaSmem *>(ba_smem_raw);
    const int nS = B.n_pairs * 36, n6 = 6 * B.n_free;
    BaPoolEntry *pool = reinterpret_cast<BaPoolEntry *>(ba_smem_raw + ((sizeof(BaSmem) + 15) & ~(size_t)15));
    double *Ssm = reinterpret_cast<double *>(pool + B.pool_cap);
    double *gsm = Ssm + nS, *bpsm = gsm + n6;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool init_mode = ST.phase == kBaNeedInit;  // only max diag(H) is wanted
    const double lambda = init_mode ? 0.0 : ST.lambda;
    const bool robust = ST.robust != 0;
    const int cur = ST.cur;
    const Pose *poses = B.poses[cur];
    const double *pts = B.pts[cur], *lines = B.lines[cur];
    const se3::Cam cam{B.fx, B.fy, B.cx, B.cy, B.bf};
    if (!kLarge)
        for (int i = tid; i < nS + 2 * n6; i += kBaThreads) Ssm[i] = 0.0;
    double chi_acc = 0.0, maxdiag = 0.0;
    const int lm_begin = B.cta_ranges[blockIdx.x], lm_end = B.cta_ranges[blockIdx.x + 1];
    const int LB = B.batch_landmarks;  // landmarks per batch (<= kBaWarps), LB * max_free_degree <= B.pool_cap
    // phase 2 ownership: thread t < 504 owns entry (r, c) = (t % 36) of the blocks p = t / 36, t / 36 + 14, ...
    const int own_rc = tid % 36, own_r = own_rc / 6, own_c = own_rc - own_r * 6, own_p0 = tid / 36;
    const int own_a = own_r < own_c ? own_r : own_c, own_b = own_r < own_c ? own_c : own_r;
    const int own_tri = own_a * 6 - own_a * (own_a - 1) / 2 + (own_b - own_a);
    __syncthreads();

    for (int batch0 = lm_begin; batch0 < lm_end; batch0 += LB) {
        const int lmb = warp;  // this warp's landmark inside the batch
        const int lm = batch0 + lmb;
        const bool has_lm = lmb < LB && lm < lm_end;
        // ---- slot table reset, count free active edges per landmark for the pool layout
        if (!kLarge) {
            for (int i = tid; i < kBaWarps * kBaMaxFree; i += kBaThreads) (&S.slot[0][0])[i] = -1;
            for (int i = tid; i < kBaMaxFree; i += kBaThreads) S.kfmask[i] = 0u;
        }
        int e0 = 0, e1 = 0;
        bool is_line = false;
        if (has_lm) {
            is_line = lm >= B.n_pts;
            const int *off = is_line ? B.ln_off : B.pt_off;
            const int li = is_line ? lm - B.n_pts : lm;
            e0 = off[li];
            e1 = off[li + 1];
        }
        const int *ekf = is_line ? B.ln_kf : B.pt_kf;
        const uint8_t *elevel = is_line ? B.ln_level : B.pt_level;
        int nfree = 0;
        for (int e = e0 + lane; e < e1; e += 32) nfree += (elevel[e] == 0 && B.kf_hidx[ekf[e]] >= 0) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nfree += __shfl_xor_sync(0xffffffffu, nfree, o);
        if (lane == 0) S.warp_cnt[warp] = has_lm ? nfree : 0;
        __syncthreads();
        // pool base of this warp = exclusive prefix of the per-warp counts (every warp scans the 16 counts itself)
        int pb;
        {
            const int cnt = lane < kBaWarps ? S.warp_cnt[lane] : 0;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < kBaWarps; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            pb = __shfl_sync(0xffffffffu, incl - cnt, warp);
        }
        const int pn = has_lm ? nfree : 0;
        // ---- phase 1: one warp per landmark
        if (has_lm && !is_line) {
            const int li = lm;
            const double *X = pts + 3 * (size_t)li;
            double hll[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};  // upper triangle of Hll
            bool any_active = false;
            int pool_cursor = pb;
            for (int ebase = e0; ebase < e1; ebase += 32) {
                const int e = ebase + lane;
                const bool active = e < e1 && elevel[e] == 0;
                double Jp[18], Jl[9], r[3] = {0, 0, 0}, w = 0;
                int h = -1;
                if (active) {
                    const int k = ekf[e];
                    h = B.kf_hidx[k];
                    const Pose &P = poses[k];
                    const float *obs = B.pt_obs + 3 * (size_t)e;
                    const double info = B.pt_info[e];
                    double pc[3];
                    const double chi2 = eval_pt(cam, P, X, obs, info, r, pc);
                    const bool stereo = !(obs[2] < 0);
                    // the third row of both Jacobians and r[2] are zero for a monocular observation: always 3 rows
                    se3::point_jac_pose(cam, pc, stereo, Jp);
                    se3::point_jac_landmark(cam, P.R, pc, stereo, Jl);
                    w = info;
                    B.pt_chi2[e] = chi2;
                    double rho0 = chi2, rho1 = 1.0;
                    if (robust) se3::huber(chi2, B.delta_pt, rho0, rho1);
                    chi_acc += rho0;
                    w *= rho1;
                    int q = 0;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
#pragma unroll
                        for (int c = a; c < 3; ++c) {
                            double s = 0;
#pragma unroll
                            for (int rr = 0; rr < 3; ++rr) s += Jl[rr * 3 + a] * w * Jl[rr * 3 + c];
                            hll[q++] += s;
                        }
                        double s = 0;
#pragma unroll
                        for (int rr = 0; rr < 3; ++rr) s += Jl[rr * 3 + a] * (-w * r[rr]);
                        bl[a] += s;
                    }
                }
                any_active = any_active || __any_sync(0xffffffffu, active);
                // free-keyframe edges get a pool entry (stable order = edge order)
                const bool freee = active && h >= 0;
                const unsigned bal = __ballot_sync(0xffffffffu, freee);
                if (freee) {
                    const int pi = pool_cursor + __popc(bal & ((1u << lane) - 1));
                    BaPoolEntry &pe = pool[pi];
                    double *Wg = B.pt_W + 24 * (size_t)e;  // W is needed again by the back-substitution
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            double s = 0;
#pragma unroll
                            for (int rr = 0; rr < 3; ++rr) s += Jp[rr * 6 + a] * w * Jl[rr * 3 + c];
                            pe.W[a * 3 + c] = s;
                            Wg[a * 3 + c] = s;
                        }
                        double s = 0;
#pragma unroll
                        for (int rr = 0; rr < 3; ++rr) s += Jp[rr * 6 + a] * (-w * r[rr]);
                        pe.bpe[a] = s;
                    }
                    int q = 0;
#pragma unroll
                    for (int a = 0; a < 6; ++a)
#pragma unroll
                        for (int c = a; c < 6; ++c) {
                            double s = 0;
#pragma unroll
                            for (int rr = 0; rr < 3; ++rr) s += Jp[rr * 6 + a] * w * Jp[rr * 6 + c];
                            pe.A[q++] = s;
                        }
                    pe.h = h;
                    if (!kLarge) {
                        S.slot[lmb][h] = (short)pi;
                        atomicOr(&S.kfmask[h], 1u << lmb);
                    }
                }
                pool_cursor += __popc(bal);
            }
            // plane edge (unary, numeric Jacobian; Huber delta = 1 in both phases)
            if (lane == 0) {
                const int pe_i = B.pt_plane ? B.pt_plane[li] : -1;
                if (pe_i >= 0) {
                    const double *fn = B.pl_fn + 4 * (size_t)pe_i;
                    const double err = eval_plane(X, fn);
                    double Jn[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        double Xp[3] = {X[0], X[1], X[2]}, Xm[3] = {X[0], X[1], X[2]};
                        Xp[d] += kDelta;
                        Xm[d] -= kDelta;
                        Jn[d] = (1.0 / (2 * kDelta)) * (eval_plane(Xp, fn) - eval_plane(Xm, fn));
                    }
                    double rho0, rho1;
                    se3::huber(err * err, 1.0, rho0, rho1);
                    chi_acc += rho0;
                    B.pl_err[pe_i] = err;
                    int q = 0;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
#pragma unroll
                        for (int c = a; c < 3; ++c) hll[q++] += Jn[a] * rho1 * Jn[c];
                        bl[a] += Jn[a] * (-rho1 * err);
                    }
                    any_active = true;
                }
            }
            any_active = __any_sync(0xffffffffu, any_active);
#pragma unroll
            for (int q = 0; q < 6; ++q) hll[q] = warp_sum(hll[q]);
#pragma unroll
            for (int a = 0; a < 3; ++a) bl[a] = warp_sum(bl[a]);
            if (lane == 0) finish_landmark<3>(B, S, warp, false, li, hll, bl, any_active, init_mode, lambda, maxdiag);
            __syncwarp();
            pool_products<3>(S, pool, warp, lane, pb, pn);
        } else if (has_lm) {
            const int li = lm - B.n_pts;
            const double *Lm = lines + 6 * (size_t)li;
            if (lane < 8) {  // perturbed lines for the numeric Jacobian w.r.t. the line vertex
                double v[4] = {0, 0, 0, 0};
                v[lane >> 1] = (lane & 1) ? -kDelta : kDelta;
                line_oplus(Lm, v, S.pert_line[warp][lane]);
            }
            __syncwarp();
            BaLineScratch &sc = S.scr[warp];
            // lanes 0..9 own the upper triangle of Hll, lanes 10..13 own bl; (oa, oc) = this lane's entry
            int oa = 0, oc = 0;
            if (lane < 10) {
                int q = lane;
                while (q >= 4 - oa) {
                    q -= 4 - oa;
                    ++oa;
                }
                oc = oa + q;
            } else if (lane < 14) {
                oa = lane - 10;
            }
            // second-round ownership of A (upper triangle of the 6 x 6 pose block), lanes 0..20
            int aa = 0, ac = 0;
            {
                int q = lane < 21 ? lane : 0;
                while (q >= 6 - aa) {
                    q -= 6 - aa;
                    ++aa;
                }
                ac = aa + q;
            }
            double hacc = 0.0;
            bool any_active = false;
            int pool_cursor = pb;
            const double scalar = 1.0 / (2 * kDelta);
            const double *Lv = lane >= 13 && lane < 21 ? S.pert_line[warp][lane - 13] : Lm;
            for (int e = e0; e < e1; ++e) {
                if (elevel[e] != 0) continue;  // warp-uniform
                any_active = true;
                const int k = ekf[e];
                const int h = B.kf_hidx[k];
                const float *obs = B.ln_obs + 4 * (size_t)e;
                const double info = B.ln_info[e];
                double ev[2] = {0, 0};
                if (lane < 21) {
                    const Pose &P = (lane >= 1 && lane < 13) ? B.pert_pose[12 * (size_t)k + (lane - 1)] : poses[k];
                    eval_ln(cam, P, Lv, obs, info, ev);
                }
                // central differences: lane 1 + 2d (pose direction d) and lane 13 + 2d (line direction d) pair with the next lane
                const double em0 = __shfl_down_sync(0xffffffffu, ev[0], 1), em1 = __shfl_down_sync(0xffffffffu, ev[1], 1);
                if (lane == 0) {
                    const double chi2 = ev[0] * (info * ev[0]) + ev[1] * (info * ev[1]);
                    B.ln_chi2[e] = chi2;
                    double rho0 = chi2, rho1 = 1.0;
                    if (robust) se3::huber(chi2, B.delta_ln, rho0, rho1);
                    chi_acc += rho0;
                    sc.r[0] = ev[0];
                    sc.r[1] = ev[1];
                    sc.w = info * rho1;
                } else if (lane < 13) {
                    if (lane & 1) {
                        const int d = (lane - 1) >> 1;
                        sc.jp[d] = scalar * (ev[0] - em0);
                        sc.jp[6 + d] = scalar * (ev[1] - em1);
                    }
                } else if (lane < 21) {
                    if (lane & 1) {
                        const int d = (lane - 13) >> 1;
                        sc.jl[d] = scalar * (ev[0] - em0);
                        sc.jl[4 + d] = scalar * (ev[1] - em1);
                    }
                }
                __syncwarp();
                const double w = sc.w, r0 = sc.r[0], r1 = sc.r[1];
                if (lane < 10) {
                    hacc += sc.jl[oa] * w * sc.jl[oc] + sc.jl[4 + oa] * w * sc.jl[4 + oc];
                } else if (lane < 14) {
                    hacc += sc.jl[oa] * (-w * r0) + sc.jl[4 + oa] * (-w * r1);
                }
                if (h >= 0) {  // warp-uniform: pool entry of a free-keyframe edge
                    BaPoolEntry &pe = pool[pool_cursor];
                    if (lane < 24) {
                        const int a = lane >> 2, c = lane & 3;
                        const double s = sc.jp[a] * w * sc.jl[c] + sc.jp[6 + a] * w * sc.jl[4 + c];
                        pe.W[lane] = s;
                        B.ln_W[24 * (size_t)e + lane] = s;  // W is needed again by the back-substitution
                    } else if (lane < 30) {
                        const int a = lane - 24;
                        pe.bpe[a] = sc.jp[a] * (-w * r0) + sc.jp[6 + a] * (-w * r1);
                    } else if (lane == 30) {
                        pe.h = h;
                        if (!kLarge) {
                            S.slot[lmb][h] = (short)pool_cursor;
                            atomicOr(&S.kfmask[h], 1u << lmb);
                        }
                    }
                    if (lane < 21) pe.A[lane] = sc.jp[aa] * w * sc.jp[ac] + sc.jp[6 + aa] * w * sc.jp[6 + ac];
                    ++pool_cursor;
                }
                __syncwarp();  // the scratch is rewritten by the next edge
            }
            if (lane < 10) S.hll[warp][lane] = hacc;
            else if (lane < 14) S.bl[warp][lane - 10] = hacc;
            __syncwarp();
            if (lane == 0) {
                double hll[10], bl[4];
#pragma unroll
                for (int q = 0; q < 10; ++q) hll[q] = S.hll[warp][q];
#pragma unroll
                for (int a = 0; a < 4; ++a) bl[a] = S.bl[warp][a];
                finish_landmark<4>(B, S, warp, true, li, hll, bl, any_active, init_mode, lambda, maxdiag);
            }
            __syncwarp();
            pool_products<4>(S, pool, warp, lane, pb, pn);
        }
        __syncthreads();
        if (kLarge) {
            // ---- phase 2 (large): the warp of a landmark adds the blocks of every pair of its free observers to the dense
            // reduced system in HBM: S(hi, hj) -= Y_i W_j^T for hi <= hj, S(h, h) += A, g(h) += gpe, bp(h) += bpe
            if (has_lm) {
                const int Dl = is_line ? 4 : 3;
                const int N = B.n_free;
                for (int w = lane; w < pn * pn * 36; w += 32) {
                    const int rc = w % 36, ij = w / 36, i = ij / pn, j = ij - i * pn;
                    const BaPoolEntry &pi = pool[pb + i], &pj = pool[pb + j];
                    if (pi.h > pj.h) continue;  // upper block triangle only (a keyframe observes a landmark once: pi.h == pj.h <=> i == j)
                    const int r = rc / 6, c = rc - r * 6;
                    double sacc = 0;
                    for (int q = 0; q < Dl; ++q) sacc += pi.Y[r * Dl + q] * pj.W[c * Dl + q];
                    sacc = -sacc;
                    if (i == j) {
                        const int a2 = r < c ? r : c, b2 = r < c ? c : r;
                        sacc += pi.A[a2 * 6 - a2 * (a2 - 1) / 2 + (b2 - a2)];
                    }
                    const size_t p = (size_t)pi.h * N - (size_t)pi.h * (pi.h - 1) / 2 + (pj.h - pi.h);
                    atomicAdd(&B.packed[p * 36 + rc], sacc);
                }
                for (int w = lane; w < pn * 6; w += 32) {
                    const int i = w / 6, r = w - i * 6;
                    const BaPoolEntry &pi = pool[pb + i];
                    atomicAdd(&B.packed[nS + 6 * pi.h + r], pi.gpe[r]);
                    atomicAdd(&B.packed[nS + n6 + 6 * pi.h + r], pi.bpe[r]);
                }
            }
            __syncthreads();
            continue;
        }
        // ---- phase 2: every thread owns fixed entries of S / g / bp (no atomics, fixed summation order)
        // only the landmarks that observe BOTH keyframes of a block contribute: walk the set bits of the two masks
        // (ascending landmark = the summation order of a dense scan)
        if (tid < 504) {
            const int n_line0 = B.n_pts - batch0;  // landmarks lb >= n_line0 of this batch are lines
            // block p = (bi, bj), bi <= bj, row-wise; the walk in steps of 14 is tracked without the index tables
            const int N = B.n_free;
            int bi = 0, bj = own_p0;
            while (bj >= N && bi < N) {  // bi == N: past the last block (the loop below does not run)
                ++bi;
                bj = bj - N + bi;
            }
            for (int p = own_p0; p < B.n_pairs; p += 14) {
                const int cbi = bi, cbj = bj;
                bj += 14;
                while (bj >= N && bi < N) {
                    ++bi;
                    bj = bj - N + bi;
                }
                unsigned m = S.kfmask[cbi] & S.kfmask[cbj];
                if (!m) continue;
                double acc = 0.0;
                while (m) {
                    const int lb = __ffs(m) - 1;
                    m &= m - 1;
                    const BaPoolEntry &pi = pool[S.slot[lb][cbi]], &pj = pool[S.slot[lb][cbj]];
                    double s;
                    if (lb >= n_line0) {
                        const double *y = pi.Y + own_r * 4, *wv = pj.W + own_c * 4;
                        s = y[0] * wv[0] + y[1] * wv[1] + y[2] * wv[2] + y[3] * wv[3];
                    } else {
                        const double *y = pi.Y + own_r * 3, *wv = pj.W + own_c * 3;
                        s = y[0] * wv[0] + y[1] * wv[1] + y[2] * wv[2];
                    }
                    acc -= s;
                    if (cbi == cbj) acc += pi.A[own_tri];
                }
                Ssm[p * 36 + own_rc] += acc;
            }
        }
        for (int ent = tid; ent < n6; ent += kBaThreads) {
            const int bi = ent / 6, r = ent - bi * 6;
            unsigned m = S.kfmask[bi];
            if (!m) continue;
            double ga = 0.0, ba = 0.0;
            while (m) {
                const int lb = __ffs(m) - 1;
                m &= m - 1;
                const int si = S.slot[lb][bi];
                ga += pool[si].gpe[r];
                ba += pool[si].bpe[r];
            }
            gsm[ent] += ga;
            bpsm[ent] += ba;
        }
        __syncthreads();
    }
    // ---- per-CTA partial system + chi2 / max-diag
    double *out = B.partial + (size_t)blockIdx.x * B.packed_len;
    if (!kLarge)
        for (int i = tid; i < nS + 2 * n6; i += kBaThreads) out[i] = Ssm[i];
    chi_acc = warp_sum(chi_acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxdiag = fmax(maxdiag, __shfl_xor_sync(0xffffffffu, maxdiag, o));
    if (lane == 0) {
        S.red[warp][0] = chi_acc;
        S.red[warp][1] = maxdiag;
    }
    __syncthreads();
    if (tid == 0) {
        double c = 0, m = 0;
        for (int w = 0; w < kBaWarps; ++w) {
            c += S.red[w][0];
            m = fmax(m, S.red[w][1]);
        }
        if (kLarge) {  // chi2 sums, max diag(H) of the landmark blocks is a maximum: both straight into the packed vector
            atomicAdd(&B.packed[nS + 2 * n6], c);
            atomicMax(reinterpret_cast<unsigned long long *>(&B.packed[nS + 2 * n6 + 1 + B.rank]),
                      (unsigned long long)__double_as_longlong(m));  // non-negative doubles order like their bit patterns
        } else {
            out[nS + 2 * n6] = c;
            out[nS + 2 * n6 + 1] = m;
        }
    }
}

// =========================================================================================================
// ba_reduce_kernel: packed = sum over CTAs of the partial systems; max-diag goes to this rank's one-hot slot
// =========================================================================================================
constexpr int kReduceLanes = 8;  // threads per packed entry: enough loads in flight to stream the partials out of L2
__global__ void __launch_bounds__(256) ba_reduce_kernel(BaDev B) {
    if (B.state->phase == kBaDone) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t / kReduceLanes, j = t % kReduceLanes;  // entry, slice of the CTA partials
    const int n_sum = B.n_pairs * 36 + 12 * B.n_free + 1;
    const int G = B.num_ctas;
    const size_t len = (size_t)B.packed_len;
    // fixed summation order: slice j adds the partials of CTA j, j + 8, ..., then the slices are combined by a shuffle tree
    double s = 0;
    if (i <= n_sum) {
        const double *src = B.partial + i;
        int g = j;
        if (i < n_sum) {
            for (; g + 3 * kReduceLanes < G; g += 4 * kReduceLanes) {
                const double v0 = __ldcg(src + (size_t)g * len), v1 = __ldcg(src + (size_t)(g + kReduceLanes) * len);
                const double v2 = __ldcg(src + (size_t)(g + 2 * kReduceLanes) * len);
                const double v3 = __ldcg(src + (size_t)(g + 3 * kReduceLanes) * len);
                s += v0;
                s += v1;
                s += v2;
                s += v3;
            }
            for (; g < G; g += kReduceLanes) s += __ldcg(src + (size_t)g * len);
        } else {
            for (; g < G; g += kReduceLanes) s = fmax(s, __ldcg(src + (size_t)g * len));
        }
    }
#pragma unroll
    for (int o = kReduceLanes / 2; o > 0; o >>= 1) {
        const double v = __shfl_xor_sync(0xffffffffu, s, o);
        s = i < n_sum ? s + v : fmax(s, v);
    }
    if (j != 0) return;
    if (i < n_sum) {
        B.packed[i] = s;
    } else if (i == n_sum) {
        // max diag of the pose blocks = diagonal of the diagonal S blocks in init mode (no Schur term yet)
        for (int w = 0; w < B.world; ++w) B.packed[n_sum + w] = (w == B.rank) ? s : 0.0;
    }
}

// =========================================================================================================
// ba_solve_kernel: 6N x 6N blocked Cholesky in shared memory (packed lower triangle)
// =========================================================================================================
constexpr int kSolveThreads = 512;
__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // j <= i

__global__ void __launch_bounds__(kSolveThreads, 1) ba_solve_kernel(BaDev B) {
#ifdef PLP_CTA_EMU
    double *solve_smem = reinterpret_cast<double *>(emu_dynamic_smem);
#else
    extern __shared__ __align__(16) double solve_smem[];
#endif
    BaState &ST = *B.state;
    const int tid = threadIdx.x;
    // Thread 0 rewrites ST.phase below (lambda initialisation) while the other warps may not have read it yet: every thread
    // must act on the SAME value, so it is read once, behind a barrier (found by running this kernel under tests/cta_emu,
    // where a late thread saw the new phase, skipped the early return and waited at a barrier nobody else reached).
    __shared__ int s_phase;
    if (tid == 0) s_phase = ST.phase;
    __syncthreads();
    const int phase = s_phase;
    if (phase == kBaDone) return;
    const int N = B.n_free, n = 6 * N, nS = B.n_pairs * 36;
    const double *packed = B.packed;
    if (phase == kBaNeedInit) {  // computeLambdaInit: tau * max diag over every active vertex
        if (tid == 0) {
            double md = 0;
            for (int w = 0; w < B.world; ++w) md = fmax(md, packed[nS + 2 * n + 1 + w]);
            for (int p = 0; p < B.n_pairs; ++p)
                if (B.pair_bi[p] == B.pair_bj[p])
                    for (int a = 0; a < 6; ++a) md = fmax(md, fabs(packed[p * 36 + a * 7]));
            ST.lambda = 1e-5 * md;
            ST.ni = 2;
            ST.phase = kBaRunning;
            ST.iter_start = 1;
            ST.have_trial = 0;
        }
        return;
    }
    double *L = solve_smem;               // n(n+1)/2
    double *rhs = L + (size_t)n * (n + 1) / 2;  // n
    double *x = rhs + n;                  // n
    double *Ld = x + n;                   // N x 21: the factored diagonal blocks (lower triangle row-wise, 1 / l_cc on the diagonal)
    __shared__ int s_ok;
    const double lambda = ST.lambda;
    if (tid < 504) {  // thread t owns entry (r, c) = t % 36 of the blocks t / 36, t / 36 + 14, ...
        const int rc = tid % 36, r = rc / 6, c = rc - r * 6;
        for (int pr = tid / 36; pr < B.n_pairs; pr += 14) {
            const int bi = B.pair_bi[pr], bj = B.pair_bj[pr];
            const int gi = bi * 6 + r, gj = bj * 6 + c;
            double v = packed[pr * 36 + rc];
            if (bi == bj) {
                if (c > r) continue;  // lower part of the (symmetric) diagonal block
                if (r == c) v += lambda;
                L[tri(gi, gj)] = v;
            } else {
                L[tri(gj, gi)] = v;  // bi < bj: entry (gi, gj) of the upper part -> (gj, gi) of the lower part
            }
        }
    }
    for (int i = tid; i < n; i += kSolveThreads) rhs[i] = packed[nS + i];
    if (tid == 0) s_ok = 1;
    __syncthreads();
    // Blocked Cholesky (L L^T) on the packed lower triangle, block = one keyframe (6 x 6), right-looking, with the
    // right-hand side carried along as an extra row "n" so that the forward substitution z = L^-1 b falls out of the
    // factorisation.  Per block column two phases / two barriers: (1) every thread that owns a row below the diagonal
    // block factors that 6 x 6 block ITSELF, in registers (the same 21 shared-memory words for everybody: broadcast
    // reads, no serial section, no barrier between factor and use) and solves its row against it; (2) the rank-6
    // update of the trailing triangle, one 4 x 4 register tile per thread.  The factored diagonal blocks go to
    // Ld (the unfactored ones stay in L: other threads may still be reading them).
    const int lane = tid & 31, warp = tid >> 5;
    for (int K = 0; K < n; K += 6) {
        const int row = K + 6 + tid;
        if (row <= n) {
            double l[21];  // lower triangle row-wise: (r, c) -> r (r + 1) / 2 + c
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) l[r * (r + 1) / 2 + c] = L[tri(K + r, K + c)];
            bool pd = true;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                double d = l[c * (c + 1) / 2 + c];
#pragma unroll
                for (int m = 0; m < c; ++m) d -= l[c * (c + 1) / 2 + m] * l[c * (c + 1) / 2 + m];
                pd = pd && (d > 0.0) && isfinite(d);  // not positive definite otherwise
                const double inv = rsqrt(d);          // the diagonal keeps 1 / l_cc
                l[c * (c + 1) / 2 + c] = inv;
#pragma unroll
                for (int r = c + 1; r < 6; ++r) {
                    double v = l[r * (r + 1) / 2 + c];
#pragma unroll
                    for (int m = 0; m < c; ++m) v -= l[r * (r + 1) / 2 + m] * l[c * (c + 1) / 2 + m];
                    l[r * (r + 1) / 2 + c] = v * inv;
                }
            }
            if (tid == 0) {
#pragma unroll
                for (int q = 0; q < 21; ++q) Ld[(K / 6) * 21 + q] = l[q];
                if (!pd) s_ok = 0;
            }
            double *a = row < n ? &L[tri(row, K)] : &rhs[K];
            double xv[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                double v = a[c];
#pragma unroll
                for (int m = 0; m < c; ++m) v -= xv[m] * l[c * (c + 1) / 2 + m];
                xv[c] = v * l[c * (c + 1) / 2 + c];
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) a[c] = xv[c];
        }
        __syncthreads();
        if (!s_ok) break;  // uniform
        // rank-6 update of the trailing lower triangle, register-tiled: one thread = one 4 x 4 tile (24 + 24 loads for 96
        // multiply-adds, all independent), then the right-hand side row
        const int m = n - K - 6;
        if (m > 0) {
            const int T = (m + 3) >> 2, ntiles = T * (T + 1) / 2;
            for (int q = tid; q < ntiles; q += kSolveThreads) {
                int ti = (int)((sqrtf(8.f * (float)q + 1.f) - 1.f) * 0.5f);
                while (ti * (ti + 1) / 2 > q) --ti;
                while ((ti + 1) * (ti + 2) / 2 <= q) ++ti;
                const int tk = q - ti * (ti + 1) / 2;
                const int i0 = K + 6 + 4 * ti, k0 = K + 6 + 4 * tk;
                int ri[4], rk[4];  // row starts in the packed triangle (rows past the end are clamped: loaded, never stored)
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    ri[a] = tri(min(i0 + a, n - 1), 0);
                    rk[a] = tri(min(k0 + a, n - 1), 0);
                }
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b2 = 0; b2 < 4; ++b2) acc[a][b2] = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    double rv[4], cv[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        rv[a] = L[ri[a] + K + c];
                        cv[a] = L[rk[a] + K + c];
                    }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b2 = 0; b2 < 4; ++b2) acc[a][b2] += rv[a] * cv[b2];
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b2 = 0; b2 < 4; ++b2) {
                        const int i = i0 + a, k = k0 + b2;
                        if (i < n && k <= i) L[ri[a] + k] -= acc[a][b2];
                    }
            }
            const double a0 = rhs[K], a1 = rhs[K + 1], a2 = rhs[K + 2], a3 = rhs[K + 3], a4 = rhs[K + 4], a5 = rhs[K + 5];
            for (int k = K + 6 + tid; k < n; k += kSolveThreads) {
                const double *lk = &L[tri(k, K)];
                rhs[k] -= a0 * lk[0] + a1 * lk[1] + a2 * lk[2] + a3 * lk[3] + a4 * lk[4] + a5 * lk[5];
            }
        }
        __syncthreads();
    }
    const int ok = s_ok;
    // rhs now holds z = L^-1 b; back substitution L^T x = z by one warp, one keyframe block per step: every lane solves
    // the 6 x 6 transposed triangle itself (registers), then the lanes subtract the block's columns from the rows above
    if (tid < 32 && ok) {
        for (int J = N - 1; J >= 0; --J) {
            double l[21], xb[6];
#pragma unroll
            for (int q = 0; q < 21; ++q) l[q] = Ld[J * 21 + q];
#pragma unroll
            for (int c = 5; c >= 0; --c) {
                double v = rhs[6 * J + c];
#pragma unroll
                for (int m = c + 1; m < 6; ++m) v -= l[m * (m + 1) / 2 + c] * xb[m];
                xb[c] = v * l[c * (c + 1) / 2 + c];
            }
            __syncwarp();
            if (tid == 0) {
#pragma unroll
                for (int c = 0; c < 6; ++c) x[6 * J + c] = xb[c];
            }
            for (int i = tid; i < 6 * J; i += 32) {
                double v = rhs[i];
#pragma unroll
                for (int c = 0; c < 6; ++c) v -= L[tri(6 * J + c, i)] * xb[c];
                rhs[i] = v;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (!ok)
        for (int i = tid; i < n; i += kSolveThreads) x[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < n; i += kSolveThreads) B.dp[i] = x[i];
    // trial poses, perturbed trial poses are produced when (if) the step is accepted
    const int cur = ST.cur;
    for (int k = tid; k < B.n_kf; k += kSolveThreads) {
        const int h = B.kf_hidx[k];
        Pose P = B.poses[cur][k];
        if (h >= 0 && ok) P = se3::oplus(P, x + 6 * h);
        B.poses[cur ^ 1][k] = P;
    }
    __syncthreads();
    if (tid == 0) {
        if (ST.iter_start) {  // currentChi = activeRobustChi2() at the start of an iteration
            ST.current_chi = packed[nS + 2 * n];
            ST.qmax = 0;
            ST.iter_start = 0;
        }
        ST.ok2 = ok;
        ST.have_trial = 1;
    }
    if (warp == 0) {  // computeScale: dp . (lambda dp + b), lane-strided then the shuffle tree
        double sc = 0;
        for (int i = lane; i < n; i += 32) sc += x[i] * (lambda * x[i] + packed[nS + n + i]);
        sc = warp_sum(sc);
        if (lane == 0) ST.scale_pose = sc;
    }
}

// =========================================================================================================
// ba_update_kernel: back-substitution, trial landmarks, trial errors
// =========================================================================================================
__global__ void __launch_bounds__(kBaThreads, 1) ba_update_kernel(BaDev B) {
    const BaState &ST = *B.state;
    if (ST.phase == kBaDone || !ST.have_trial) return;
#ifdef PLP_CTA_EMU
    double *s_dp = reinterpret_cast<double *>(emu_dynamic_smem);  // 6 x n_free
#else
    extern __shared__ __align__(16) double s_dp[];  // 6 x n_free
#endif
    __shared__ double s_red[kBaWarps][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cur = ST.cur, nxt = cur ^ 1;
    const double lambda = ST.lambda;
    const bool robust = ST.robust != 0, ok2 = ST.ok2 != 0;
    const se3::Cam cam{B.fx, B.fy, B.cx, B.cy, B.bf};
    for (int i = tid; i < 6 * B.n_free; i += kBaThreads) s_dp[i] = B.dp[i];
    __syncthreads();
    const Pose *tposes = B.poses[nxt];
    double chi_acc = 0.0, scale_acc = 0.0;
    const int lm_begin = B.cta_ranges[blockIdx.x], lm_end = B.cta_ranges[blockIdx.x + 1];
    for (int lm = lm_begin + warp; lm < lm_end; lm += kBaWarps) {
        const bool is_line = lm >= B.n_pts;
        const int D = is_line ? 4 : 3;
        const int li = is_line ? lm - B.n_pts : lm;
        const int *off = is_line ? B.ln_off : B.pt_off;
        const int e0 = off[li], e1 = off[li + 1];
        const int *ekf = is_line ? B.ln_kf : B.pt_kf;
        const uint8_t *elevel = is_line ? B.ln_level : B.pt_level;
        const bool act = (is_line ? B.ln_active : B.pt_active)[li] != 0;
        // cl = bl - sum_e W_e^T dp_h(e)
        double cl[4] = {0, 0, 0, 0};
        if (act && ok2) {
            for (int e = e0 + lane; e < e1; e += 32) {
                if (elevel[e]) continue;
                const int h = B.kf_hidx[ekf[e]];
                if (h < 0) continue;
                const double *W = (is_line ? B.ln_W : B.pt_W) + 24 * (size_t)e;
                for (int c = 0; c < D; ++c) {
                    double s = 0;
                    for (int a = 0; a < 6; ++a) s += W[a * D + c] * s_dp[6 * h + a];
                    cl[c] -= s;
                }
            }
        }
        for (int c = 0; c < D; ++c) cl[c] = warp_sum(cl[c]);
        double dl[4] = {0, 0, 0, 0};
        const double *bl = (is_line ? B.ln_bl + 4 * (size_t)li : B.pt_bl + 4 * (size_t)li);
        if (act && ok2) {
            const double *Di = (is_line ? B.ln_Dinv + 16 * (size_t)li : B.pt_Dinv + 16 * (size_t)li);
            for (int a = 0; a < D; ++a) {
                double s = 0;
                for (int c = 0; c < D; ++c) s += Di[a * D + c] * (bl[c] + cl[c]);
                dl[a] = s;
            }
            if (lane == 0)
                for (int a = 0; a < D; ++a) scale_acc += dl[a] * (lambda * dl[a] + bl[a]);
        }
        // trial landmark
        double Xt[6];
        if (!is_line) {
            const double *X = B.pts[cur] + 3 * (size_t)li;
            for (int a = 0; a < 3; ++a) Xt[a] = X[a] + dl[a];
            if (lane == 0)
                for (int a = 0; a < 3; ++a) B.pts[nxt][3 * (size_t)li + a] = Xt[a];
        } else {
            const double *Lc = B.lines[cur] + 6 * (size_t)li;
            if (act && ok2)
                line_oplus(Lc, dl, Xt);
            else
                for (int a = 0; a < 6; ++a) Xt[a] = Lc[a];
            if (lane == 0)
                for (int a = 0; a < 6; ++a) B.lines[nxt][6 * (size_t)li + a] = Xt[a];
        }
        // errors at the trial state for the active edges (they stay even if the step is rejected)
        for (int e = e0 + lane; e < e1; e += 32) {
            if (elevel[e]) continue;
            const Pose &P = tposes[ekf[e]];
            double r[3], chi2;
            if (!is_line) {
                double pc[3];
                chi2 = eval_pt(cam, P, Xt, B.pt_obs + 3 * (size_t)e, B.pt_info[e], r, pc);
                B.pt_chi2[e] = chi2;
                double rho0 = chi2, rho1;
                if (robust) se3::huber(chi2, B.delta_pt, rho0, rho1);
                chi_acc += rho0;
            } else {
                chi2 = eval_ln(cam, P, Xt, B.ln_obs + 4 * (size_t)e, B.ln_info[e], r);
                B.ln_chi2[e] = chi2;
                double rho0 = chi2, rho1;
                if (robust) se3::huber(chi2, B.delta_ln, rho0, rho1);
                chi_acc += rho0;
            }
        }
        if (!is_line && lane == 0 && B.pt_plane) {
            const int pe_i = B.pt_plane[li];
            if (pe_i >= 0) {
                const double err = eval_plane(Xt, B.pl_fn + 4 * (size_t)pe_i);
                B.pl_err[pe_i] = err;
                double rho0, rho1;
                se3::huber(err * err, 1.0, rho0, rho1);
                chi_acc += rho0;
            }
        }
    }
    chi_acc = warp_sum(chi_acc);
    scale_acc = warp_sum(scale_acc);
    if (lane == 0) {
        s_red[warp][0] = chi_acc;
        s_red[warp][1] = scale_acc;
    }
    __syncthreads();
    __shared__ int s_last;
    unsigned *done = reinterpret_cast<unsigned *>(B.trial_sum + 4);  // zero between kernels (reset by the last CTA)
    if (tid == 0) {
        double c = 0, s = 0;
        for (int w = 0; w < kBaWarps; ++w) {
            c += s_red[w][0];
            s += s_red[w][1];
        }
        B.trial_partial[2 * blockIdx.x] = c;
        B.trial_partial[2 * blockIdx.x + 1] = s;
        __threadfence();
        s_last = atomicAdd(done, 1u) == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    // the CTA that finishes last sums the per-CTA partials in a fixed order (lane-strided, then the shuffle tree);
    // multi-GPU runs all-reduce trial_sum[0..1] over the ranks afterwards
    if (s_last && warp == 0) {
        __threadfence();
        double c = 0, s = 0;
        for (int g = lane; g < (int)gridDim.x; g += 32) {
            c += __ldcg(B.trial_partial + 2 * g);
            s += __ldcg(B.trial_partial + 2 * g + 1);
        }
        c = warp_sum(c);
        s = warp_sum(s);
        if (lane == 0) {
            B.trial_sum[0] = c;
            B.trial_sum[1] = s;
            *done = 0u;
        }
    }
}

// =========================================================================================================
// ba_decide_kernel: OptimizationAlgorithmLevenberg accept / reject + SparseOptimizer::optimize loop control
// =========================================================================================================
__global__ void __launch_bounds__(512) ba_decide_kernel(BaDev B) {
    BaState &ST = *B.state;
    // thread 0 rewrites the state below: the entry conditions are read once, behind a barrier (see ba_solve_kernel)
    __shared__ int s_entry[2];
    if (threadIdx.x == 0) {
        s_entry[0] = ST.phase;
        s_entry[1] = ST.have_trial;
    }
    __syncthreads();
    if (s_entry[0] == kBaDone) return;
    if (s_entry[1]) {
        if (threadIdx.x == 0) {
            double temp_chi = B.trial_sum[0];
            if (!ST.ok2) temp_chi = 1.7976931348623157e308;
            double rho = ST.current_chi - temp_chi;
            const double scale = ST.scale_pose + B.trial_sum[1] + 1e-3;
            rho /= scale;
            bool lambda_finite = true;
            ST.tries++;
            if (rho > 0 && isfinite(temp_chi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = fmin(alpha, 2. / 3.);
                ST.lambda *= fmax(1. / 3., alpha);
                ST.ni = 2;
                ST.current_chi = temp_chi;
                ST.cur ^= 1;  // commit the trial state
                ST.accepted = 1;
            } else {
                ST.lambda *= ST.ni;
                ST.ni *= 2;
                ST.accepted = 0;
                if (!isfinite(ST.lambda)) lambda_finite = false;
            }
            if (lambda_finite) ST.qmax++;
            ST.rho = rho;
            const bool again = lambda_finite && rho < 0 && ST.qmax < 10;
            if (!again) {  // the LM iteration is over
                ST.it++;
                ST.iter_start = 1;
                const bool terminate = (ST.qmax == 10 || rho == 0 || !lambda_finite);
                if (terminate || ST.it >= ST.max_it) ST.phase = kBaDone;
            }
            ST.have_trial = 0;
        }
    }
    __syncthreads();
    // numeric-Jacobian support: perturbed poses of the (possibly new) current estimate
    if (B.n_ln_edges > 0 && ST.phase != kBaDone) {
        const int cur = ST.cur;
        for (int i = threadIdx.x; i < B.n_kf * 12; i += blockDim.x) {
            const int k = i / 12, d = i - k * 12;
            double u[6] = {0, 0, 0, 0, 0, 0};
            u[d >> 1] = (d & 1) ? -kDelta : kDelta;
            B.pert_pose[i] = se3::oplus(B.poses[cur][k], u);
        }
    }
}

// =========================================================================================================
// classification between / after the two optimize() calls (local_bundle_adjuster.cc:303-372)
// =========================================================================================================
__global__ void ba_classify_kernel(BaDev B, int set_levels) {
    const BaState &ST = *B.state;
    const int cur = ST.cur;
    const se3::Cam cam{B.fx, B.fy, B.cx, B.cy, B.bf};
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double chi_sq_2D = (double)5.99146f, chi_sq_3D = (double)7.81473f;
    if (i < B.n_pt_edges) {
        const float *obs = B.pt_obs + 3 * (size_t)i;
        const bool stereo = !(obs[2] < 0);
        double pc[3];
        se3::map_point(B.poses[cur][B.pt_kf[i]].R, B.poses[cur][B.pt_kf[i]].t, B.pts[cur] + 3 * (size_t)B.pt_lm[i], pc);
        const bool out = (stereo ? chi_sq_3D : chi_sq_2D) < B.pt_chi2[i] || !(0.0 < pc[2]);
        if (set_levels) {
            if (out) B.pt_level[i] = 1;
        } else {
            B.pt_outlier[i] = out ? 1 : 0;
        }
    } else if (i < B.n_pt_edges + B.n_ln_edges) {
        const int e = i - B.n_pt_edges;
        const bool out = chi_sq_2D < B.ln_chi2[e] ||
                         !line_depth_positive(cam, B.poses[cur][B.ln_kf[e]], B.lines[cur] + 6 * (size_t)B.ln_lm[e],
                                              B.ln_obs + 4 * (size_t)e);
        if (set_levels) {
            if (out) B.ln_level[e] = 1;
        } else {
            B.ln_outlier[e] = out ? 1 : 0;
        }
    }
}

__global__ void ba_init_poses_kernel(BaDev B, const double *T_in) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= B.n_kf) return;
    const Pose P = se3::from_matrix(T_in + 16 * (size_t)k);
    B.poses[0][k] = P;
    B.poses[1][k] = P;
}

__global__ void ba_export_kernel(BaDev B, double *T_out, double *pts_out, double *lines_out) {
    const int cur = B.state->cur;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B.n_kf) {
        double T[16];
        se3::to_matrix(B.poses[cur][i], T);
        for (int k = 0; k < 16; ++k) T_out[16 * (size_t)i + k] = T[k];
    }
    if (i < 3 * B.n_pts) pts_out[i] = B.pts[cur][i];
    if (i < 6 * B.n_lines) lines_out[i] = B.lines[cur][i];
}

__global__ void ba_set_state_kernel(BaDev B, int max_it, int robust, int reset_cur) {
    BaState &ST = *B.state;
    ST.phase = kBaNeedInit;
    ST.it = 0;
    ST.max_it = max_it;
    ST.qmax = 0;
    ST.iter_start = 1;
    ST.have_trial = 0;
    ST.ok2 = 0;
    ST.robust = robust;
    ST.lambda = 0;
    ST.ni = 2;
    ST.rho = 0;
    ST.accepted = 0;
    ST.solve_active = 0;
    if (reset_cur) {
        ST.cur = 0;
        ST.tries = 0;
        ST.current_chi = 0;
    }
}


}  // namespace balm
}  // namespace plp
