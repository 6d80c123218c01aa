// fuse.cu -- match::fuse search kernels (sm_100a).
//
// Replaces the per-landmark search of match/fuse.cc:40-151 (detect_duplication), :153-300 (replace_duplication) and
// :304-503 (replace_duplication_line) of the reference.  Unlike the projection matchers, fuse.cc has no "skip keypoints
// claimed by an earlier landmark": every landmark's best keypoint depends only on the landmark and on the target
// keyframe's features, so a (target keyframe x landmark) batch is embarrassingly parallel -- grid = (landmark chunks,
// targets), one WARP per landmark.  Only the effects (add_observation / replace) are sequential; they stay in the adapter.
//
// Exactness: reprojection, the distance / viewing-angle gates and the chi-square gates are evaluated in the reference's
// double / float mix (this file is compiled with -fmad=false); predict_scale_level's logf becomes a comparison of the
// float ratio against thresholds derived on the host from the caller's own libm (see build_level_thresholds);
// candidates are visited in get_keypoints_in_cell order (cell-x, cell-y, insertion) via the same rank sort as match.cu,
// and "first strictly smaller distance wins" is min(distance << 32 | traversal position).
#include "common.cuh"
#include "pack.cuh"

#include <math.h>
#include <algorithm>
#include <cmath>
#include <mutex>

namespace plp {

namespace {

constexpr int kThreads = 512;
constexpr int kFuseMaxPoints = 3072;  // keypoints per target keyframe (shared-memory bound, like the window matcher)
constexpr int kFuseMaxLines = 2048;   // keylines per target keyframe
constexpr int kMaxLevels = 32;

struct FusePointTarget {
    int n;
    const float *x, *y, *xr;  // xr may be null (monocular keyframe)
    const int32_t *octave;
    const uint8_t *desc;
    const uint8_t *skip;  // may be null
    double R[9], t[3], c[3];
};

struct FuseLineTarget {
    int n;
    const float *sx, *sy, *ex, *ey;
    const int32_t *octave;
    const uint8_t *desc;
    const uint8_t *skip;  // may be null
    double R[9], t[3], c[3];
};

struct FuseLandmarks {
    int m;
    const double *pos_w;
    const double *normal;  // points only
    const float *min_d, *max_d, *max_raw;
    const uint8_t *desc;
    const uint8_t *valid;  // may be null
};

struct FuseParams {
    plp_camera cam;
    plp_grid grid;  // points only
    float scale_factors[kMaxLevels];
    float inv_sigma_sq[kMaxLevels];
    float level_thr[kMaxLevels];  // level_thr[k], 1 <= k < num_levels: smallest ratio whose predicted level is >= k
    int num_levels;
    float margin;
    int mode;
};

__device__ __forceinline__ void load_desc(const uint8_t *p, uint4 &a, uint4 &b) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    a = __ldg(q);
    b = __ldg(q + 1);
}

// data/landmark.cc:341-362 through the host-derived threshold table
__device__ __forceinline__ int predict_level(float ratio, const FuseParams &P) {
    int lvl = 0;
    for (int k = 1; k < P.num_levels; ++k) lvl += (ratio >= P.level_thr[k]) ? 1 : 0;
    return lvl;
}

struct Proj {
    double u, v;
    float x_right;
    bool in_image;
};

// camera/perspective.cc:190-209; a point behind the camera leaves (u, v) = (0, 0) (the reference leaves them unset)
__device__ __forceinline__ Proj reproject(const plp_camera &cam, const double *R, const double *t, const double *X) {
    Proj r;
    r.u = 0.0;
    r.v = 0.0;
    r.x_right = 0.0f;
    r.in_image = false;
    const double pc0 = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double pc1 = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double pc2 = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    if (pc2 <= 0.0) return r;
    const double z_inv = 1.0 / pc2;
    r.u = cam.fx * pc0 * z_inv + cam.cx;
    r.v = cam.fy * pc1 * z_inv + cam.cy;
    r.x_right = (float)(r.u - cam.focal_x_baseline * z_inv);
    r.in_image = (cam.min_x < r.u && r.u < cam.max_x && cam.min_y < r.v && r.v < cam.max_y);
    return r;
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other < k ? other : k;
    }
    return k;
}

static size_t fuse_point_smem_bytes(int cap, int cells) {
    return (size_t)cap * (32 + 6 * 4) + (size_t)(cells + 2) * 8 + 32 * 4;
}

// ---------------------------------------------------------------------------------------------------------------
// points: fuse.cc:40-151 (mode 0) / :153-300 (mode 1)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
    fuse_points_kernel(const FusePointTarget *__restrict__ targets, FuseLandmarks L, FuseParams P, int cap, int chunk,
                       int32_t *__restrict__ best_idx_out, uint16_t *__restrict__ best_dist_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const FusePointTarget &T = targets[blockIdx.y];
    const plp_grid &grid = P.grid;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kThreads / 32;
    const int cells = grid.num_cols * grid.num_rows;
    const int n = T.n;
    const int i0 = blockIdx.x * chunk, i1 = min(L.m, i0 + chunk);
    int32_t *out_idx = best_idx_out + (size_t)blockIdx.y * L.m;
    uint16_t *out_dist = best_dist_out ? best_dist_out + (size_t)blockIdx.y * L.m : nullptr;

    // carve
    uint8_t *base = smem_raw;
    uint4 *s_desc = reinterpret_cast<uint4 *>(base);
    base += (size_t)cap * 32;
    float *s_x = reinterpret_cast<float *>(base);
    base += (size_t)cap * 4;
    float *s_y = reinterpret_cast<float *>(base);
    base += (size_t)cap * 4;
    float *s_xr = reinterpret_cast<float *>(base);
    base += (size_t)cap * 4;
    int *s_oct = reinterpret_cast<int *>(base);
    base += (size_t)cap * 4;
    int *s_orig = reinterpret_cast<int *>(base);
    base += (size_t)cap * 4;
    int *s_key = reinterpret_cast<int *>(base);
    base += (size_t)cap * 4;
    int *s_cell_start = reinterpret_cast<int *>(base);
    base += (size_t)(cells + 2) * 4;
    int *s_cursor = reinterpret_cast<int *>(base);
    base += (size_t)(cells + 2) * 4;
    int *s_warp = reinterpret_cast<int *>(base);

    // ---- 1. cell key of every keypoint (data/common.h:104-109) and the cell histogram
    for (int k = tid; k < cells + 2; k += kThreads) s_cell_start[k] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kThreads) {
        const float px = T.x[i], py = T.y[i];
        const int cx = cv_floor((double)(px - grid.min_x) * grid.inv_cell_width);
        const int cy = cv_floor((double)(py - grid.min_y) * grid.inv_cell_height);
        const bool in = (0 <= cx && cx < grid.num_cols && 0 <= cy && cy < grid.num_rows);
        const int key = in ? cx * grid.num_rows + cy : cells;  // out-of-grid keypoints sort last and are never visited
        s_key[i] = key;
        atomicAdd(&s_cell_start[key], 1);
    }
    __syncthreads();
    // ---- 2. stable counting sort by (cell key, index) = traversal order of get_keypoints_in_cell
    //         (data/common.cc:275-309).  (a) exclusive scan of the histogram: cell_start[k] = first sorted position
    //         whose key >= k; (b) one warp walks the keypoints in index order, __match_any groups equal keys and gives
    //         every lane its rank inside the group, the group leader advances the cell cursor.
    {
        const int total = cells + 1;  // keys 0 .. cells
        const int per = (total + kThreads - 1) / kThreads;
        const int b0 = min(total, tid * per), b1 = min(total, b0 + per);
        int sum = 0;
        for (int k = b0; k < b1; ++k) sum += s_cell_start[k];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int v = lane < nwarps ? s_warp[lane] : 0;
            int sc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, sc, o);
                if (lane >= o) sc += u;
            }
            if (lane < nwarps) s_warp[lane] = sc - v;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        int run = s_warp[warp] + incl - sum;
        for (int k = b0; k < b1; ++k) {
            const int v = s_cell_start[k];
            s_cell_start[k] = run;
            s_cursor[k] = run;
            run += v;
        }
        if (tid == 0) s_cell_start[cells + 1] = n;
        __syncthreads();
        if (warp == 0) {
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                const int key = i < n ? s_key[i] : (0x40000000 | lane);  // idle lanes get unique keys
                const unsigned peers = __match_any_sync(0xffffffffu, key);
                const int leader = __ffs(peers) - 1;
                const int rank = __popc(peers & ((1u << lane) - 1));
                int b = 0;
                if (lane == leader && i < n) {
                    b = s_cursor[key];
                    s_cursor[key] = b + __popc(peers);
                }
                b = __shfl_sync(0xffffffffu, b, leader);
                if (i < n) s_orig[b + rank] = i;
                __syncwarp();
            }
        }
    }
    __syncthreads();
    const int n_in = s_cell_start[cells];
    // ---- 3. gather in sorted order; s_oct keeps the octave, the cell key is re-read from s_key through s_orig
    for (int p = tid; p < n_in; p += kThreads) {
        const int i = s_orig[p];
        s_x[p] = T.x[i];
        s_y[p] = T.y[i];
        s_xr[p] = T.xr ? T.xr[i] : -1.0f;
        s_oct[p] = T.octave[i];
        uint4 d0, d1;
        load_desc(T.desc + 32 * (size_t)i, d0, d1);
        s_desc[2 * p] = d0;
        s_desc[2 * p + 1] = d1;
    }
    __syncthreads();

    // ---- 4. one warp per landmark
    for (int i = i0 + warp; i < i1; i += nwarps) {
        int out = -1;
        unsigned outd = 0xFFFFu;
        bool ok = (L.valid ? L.valid[i] != 0 : true) && (T.skip ? T.skip[i] == 0 : true);
        double u = 0.0, v = 0.0;
        float q_xr = 0.0f;
        unsigned pred = 0;
        if (ok) {
            const double X[3] = {L.pos_w[3 * (size_t)i], L.pos_w[3 * (size_t)i + 1], L.pos_w[3 * (size_t)i + 2]};
            const Proj pr = reproject(P.cam, T.R, T.t, X);
            ok = pr.in_image;
            u = pr.u;
            v = pr.v;
            q_xr = pr.x_right;
            if (ok) {
                const double d0 = X[0] - T.c[0], d1 = X[1] - T.c[1], d2 = X[2] - T.c[2];
                const double dist = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                const float max_d = L.max_d[i], min_d = L.min_d[i];
                if (dist < (double)min_d || (double)max_d < dist) ok = false;
                const double *nm = L.normal + 3 * (size_t)i;
                if (ok && (d0 * nm[0] + d1 * nm[1] + d2 * nm[2] < 0.5 * dist)) ok = false;
                if (ok) pred = (unsigned)predict_level(L.max_raw[i] / (float)dist, P);
            }
        }
        if (ok && n_in > 0) {
            const float ref_x = (float)u, ref_y = (float)v, r = P.margin * P.scale_factors[pred];
            // data/common.cc:249-272
            const int min_cx = max(0, cv_floor((double)(ref_x - grid.min_x - r) * grid.inv_cell_width));
            const int max_cx = min(grid.num_cols - 1, cv_ceil((double)(ref_x - grid.min_x + r) * grid.inv_cell_width));
            const int min_cy = max(0, cv_floor((double)(ref_y - grid.min_y - r) * grid.inv_cell_height));
            const int max_cy = min(grid.num_rows - 1, cv_ceil((double)(ref_y - grid.min_y + r) * grid.inv_cell_height));
            unsigned long long best = ~0ull;
            if (min_cx < grid.num_cols && max_cx >= 0 && min_cy < grid.num_rows && max_cy >= 0) {
                uint4 q0, q1;
                load_desc(L.desc + 32 * (size_t)i, q0, q1);
                const int pred_i = (int)pred;
                for (int c = min_cx; c <= max_cx; ++c) {
                    const int p_begin = s_cell_start[c * grid.num_rows + min_cy];
                    const int p_end = s_cell_start[c * grid.num_rows + max_cy + 1];
                    for (int p = p_begin + lane; p < p_end; p += 32) {
                        const float kx = s_x[p], ky = s_y[p];
                        const float dx = kx - ref_x, dy = ky - ref_y;
                        if (!(fabsf(dx) < r && fabsf(dy) < r)) continue;  // data/common.cc:300-306
                        const int oct = s_oct[p];
                        if (P.mode == PLP_FUSE_DETECT) {
                            if (oct < pred_i - 1 || pred_i < oct) continue;  // fuse.cc:117-121 (int)
                        } else {
                            const unsigned sl = (unsigned)oct;
                            if (sl < pred - 1u || pred < sl) continue;  // fuse.cc:232-236 (unsigned, wraps at pred == 0)
                            if (sl >= (unsigned)P.num_levels) continue;  // inv_level_sigma_sq_.at() would throw
                            const double e_x = u - (double)kx, e_y = v - (double)ky;
                            const float kxr = s_xr[p];
                            if (kxr >= 0) {  // :238-251
                                const float e_xr = q_xr - kxr;
                                const double err = e_x * e_x + e_y * e_y + (double)(e_xr * e_xr);
                                if ((double)7.81473f < err * (double)P.inv_sigma_sq[sl]) continue;
                            } else {  // :252-265
                                const double err = e_x * e_x + e_y * e_y;
                                if ((double)5.99146f < err * (double)P.inv_sigma_sq[sl]) continue;
                            }
                        }
                        const unsigned d = (unsigned)hamming256(q0, q1, s_desc[2 * p], s_desc[2 * p + 1]);
                        const unsigned long long key = ((unsigned long long)d << 32) | (unsigned)p;
                        best = key < best ? key : best;
                    }
                }
            }
            best = warp_min_u64(best);
            if (best != ~0ull) {
                const unsigned d = (unsigned)(best >> 32);
                // best_dist starts at MAX_HAMMING_DIST with a strict '<' (:221-276), then HAMMING_DIST_THR_LOW (:279-282)
                if (d < (unsigned)PLP_MAX_HAMMING_DIST && d <= (unsigned)PLP_HAMMING_DIST_THR_LOW) {
                    out = s_orig[(int)(best & 0xffffffffull)];
                    outd = d;
                }
            }
        }
        if (lane == 0) {
            out_idx[i] = out;
            if (out_dist) out_dist[i] = (uint16_t)outd;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// lines: fuse.cc:304-503
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
    fuse_lines_kernel(const FuseLineTarget *__restrict__ targets, FuseLandmarks L, FuseParams P, int cap, int chunk,
                      int32_t *__restrict__ best_idx_out, uint16_t *__restrict__ best_dist_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const FuseLineTarget &T = targets[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kThreads / 32;
    const int n = T.n;
    const int i0 = blockIdx.x * chunk, i1 = min(L.m, i0 + chunk);
    int32_t *out_idx = best_idx_out + (size_t)blockIdx.y * L.m;
    uint16_t *out_dist = best_dist_out ? best_dist_out + (size_t)blockIdx.y * L.m : nullptr;

    uint8_t *base = smem_raw;
    uint4 *s_desc = reinterpret_cast<uint4 *>(base);
    base += (size_t)cap * 32;
    float *s_sx = reinterpret_cast<float *>(base);
    base += (size_t)cap * 4;
    float *s_sy = reinterpret_cast<float *>(base);
    base += (size_t)cap * 4;
    float *s_ex = reinterpret_cast<float *>(base);
    base += (size_t)cap * 4;
    float *s_ey = reinterpret_cast<float *>(base);
    base += (size_t)cap * 4;
    int *s_oct = reinterpret_cast<int *>(base);
    for (int p = tid; p < n; p += kThreads) {
        s_sx[p] = T.sx[p];
        s_sy[p] = T.sy[p];
        s_ex[p] = T.ex[p];
        s_ey[p] = T.ey[p];
        s_oct[p] = T.octave[p];
        uint4 d0, d1;
        load_desc(T.desc + 32 * (size_t)p, d0, d1);
        s_desc[2 * p] = d0;
        s_desc[2 * p + 1] = d1;
    }
    __syncthreads();

    for (int i = i0 + warp; i < i1; i += nwarps) {
        int out = -1;
        unsigned outd = 0xFFFFu;
        bool ok = (L.valid ? L.valid[i] != 0 : true) && (T.skip ? T.skip[i] == 0 : true);
        double su = 0.0, sv = 0.0, eu = 0.0, ev = 0.0;
        unsigned pred = 0;
        if (ok) {
            const double *pw = L.pos_w + 6 * (size_t)i;
            const double S[3] = {pw[0], pw[1], pw[2]}, E[3] = {pw[3], pw[4], pw[5]};
            const Proj ps = reproject(P.cam, T.R, T.t, S), pe = reproject(P.cam, T.R, T.t, E);
            su = ps.u;
            sv = ps.v;
            eu = pe.u;
            ev = pe.v;
            if (!ps.in_image && !pe.in_image) ok = false;  // :341-345
            const double M[3] = {0.5 * (S[0] + E[0]), 0.5 * (S[1] + E[1]), 0.5 * (S[2] + E[2])};
            if (ok && (!ps.in_image || !pe.in_image)) {  // :347-366
                const Proj pm = reproject(P.cam, T.R, T.t, M);
                if (!pm.in_image) ok = false;
            }
            if (ok) {  // :368-392
                const double a0 = S[0] - T.c[0], a1 = S[1] - T.c[1], a2 = S[2] - T.c[2];
                const double b0 = E[0] - T.c[0], b1 = E[1] - T.c[1], b2 = E[2] - T.c[2];
                const double dist_sp = sqrt(a0 * a0 + a1 * a1 + a2 * a2), dist_ep = sqrt(b0 * b0 + b1 * b1 + b2 * b2);
                const double min_d = (double)L.min_d[i], max_d = (double)L.max_d[i];
                if (dist_sp < min_d || max_d < dist_sp || dist_ep < min_d || max_d < dist_ep) ok = false;
                const double m0 = M[0] - T.c[0], m1 = M[1] - T.c[1], m2 = M[2] - T.c[2];
                const double dist_mp = sqrt(m0 * m0 + m1 * m1 + m2 * m2);
                if (ok) pred = (unsigned)predict_level(L.max_raw[i] / (float)dist_mp, P);
            }
        }
        if (ok && n > 0) {
            const float r = P.margin * P.scale_factors[pred];
            // data/common.cc:315-364: the candidate test builds the line from the FLOAT reprojections ...
            const double ax = (double)(float)su, ay = (double)(float)sv, bx = (double)(float)eu, by = (double)(float)ev;
            const double f0 = ay * 1.0 - 1.0 * by, f1 = 1.0 * bx - ax * 1.0, f2 = ax * by - ay * bx;
            const double fden = sqrt(f0 * f0 + f1 * f1);
            // ... and the chi-square gate from the double ones (fuse.cc:417-431)
            const double l0 = sv * 1.0 - 1.0 * ev, l1 = 1.0 * eu - su * 1.0, l2 = su * ev - sv * eu;
            const double lden = sqrt(l0 * l0 + l1 * l1);
            uint4 q0, q1;
            load_desc(L.desc + 32 * (size_t)i, q0, q1);
            unsigned long long best = ~0ull;
            for (int p = lane; p < n; p += 32) {
                const double ksx = (double)s_sx[p], ksy = (double)s_sy[p], kex = (double)s_ex[p], key_ = (double)s_ey[p];
                const float dsp = (float)((ksx * f0 + ksy * f1 + f2) / fden);
                const float dep = (float)((kex * f0 + key_ * f1 + f2) / fden);
                if (fabsf(dsp) > r || fabsf(dep) > r) continue;
                const unsigned sl = (unsigned)s_oct[p];
                if (sl >= (unsigned)P.num_levels) continue;  // _inv_level_sigma_sq_lsd.at() would throw
                const double e_sp = (ksx * l0 + ksy * l1 + l2) / lden;
                const double e_ep = (kex * l0 + key_ * l1 + l2) / lden;
                if ((double)5.99146f < (e_sp * e_sp + e_ep * e_ep) * (double)P.inv_sigma_sq[sl]) continue;
                const unsigned d = (unsigned)hamming256(q0, q1, s_desc[2 * p], s_desc[2 * p + 1]);
                const unsigned long long key = ((unsigned long long)d << 32) | (unsigned)p;
                best = key < best ? key : best;
            }
            best = warp_min_u64(best);
            if (best != ~0ull) {
                const unsigned d = (unsigned)(best >> 32);
                if (d < (unsigned)PLP_MAX_HAMMING_DIST && d <= (unsigned)PLP_HAMMING_DIST_THR_LOW) {
                    out = (int)(best & 0xffffffffull);
                    outd = d;
                }
            }
        }
        if (lane == 0) {
            out_idx[i] = out;
            if (out_dist) out_dist[i] = (uint16_t)outd;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// data/landmark.cc:349-361 as the reference evaluates it on the host
static int host_pred(float ratio, float lsf) {
    const float q = std::ceil(std::log(ratio) / lsf);
    if (!(q == q)) return INT32_MIN;  // NaN -> "negative" like the x86 conversion
    if (q >= 2147483648.0f || q < -2147483648.0f) return INT32_MIN;
    return (int)q;
}

// level_thr[k] (1 <= k < num_levels) = smallest positive float r with host_pred(r) >= k.  logf is monotone in every
// libm we know of; the neighbourhood of each threshold is checked and a violation is reported, never papered over.
struct ThresholdCacheEntry {
    float lsf;
    int num_levels;
    float thr[kMaxLevels];
};
static std::mutex g_thr_mutex;
static std::vector<ThresholdCacheEntry> g_thr_cache;  // a process sees a handful of (scale factor, levels) pairs

static plp_status build_level_thresholds_uncached(float lsf, int num_levels, float *thr);

// the table costs ~60 k logf evaluations (bisection + the monotonicity check): computed once per (lsf, num_levels)
static plp_status build_level_thresholds(float lsf, int num_levels, float *thr) {
    PLP_REQUIRE(lsf > 0.0f && num_levels >= 1 && num_levels <= kMaxLevels, "log_scale_factor / num_levels");
    std::lock_guard<std::mutex> lock(g_thr_mutex);
    for (const ThresholdCacheEntry &e : g_thr_cache)
        if (memcmp(&e.lsf, &lsf, 4) == 0 && e.num_levels == num_levels) {
            memcpy(thr, e.thr, sizeof(e.thr));
            return PLP_OK;
        }
    ThresholdCacheEntry e;
    e.lsf = lsf;
    e.num_levels = num_levels;
    PLP_TRY(build_level_thresholds_uncached(lsf, num_levels, e.thr));
    if (g_thr_cache.size() < 64) g_thr_cache.push_back(e);
    memcpy(thr, e.thr, sizeof(e.thr));
    return PLP_OK;
}

static plp_status build_level_thresholds_uncached(float lsf, int num_levels, float *thr) {
    for (int k = 0; k < kMaxLevels; ++k) thr[k] = INFINITY;
    for (int k = 1; k < num_levels; ++k) {
        uint32_t lo = 0x00000001u, hi = 0x7f7fffffu;  // predicate false at lo, true at hi
        float flo, fhi;
        memcpy(&flo, &lo, 4);
        memcpy(&fhi, &hi, 4);
        if (host_pred(flo, lsf) >= k || host_pred(fhi, lsf) < k) {
            set_error("predict_scale_level: no threshold for level %d (log_scale_factor %g)", k, (double)lsf);
            return PLP_ERR_INVALID;
        }
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            float f;
            memcpy(&f, &mid, 4);
            if (host_pred(f, lsf) >= k)
                hi = mid;
            else
                lo = mid;
        }
        for (uint32_t b = hi > 4096 ? hi - 4096 : 1; b < hi + 4096 && b <= 0x7f7fffffu; ++b) {
            float f;
            memcpy(&f, &b, 4);
            if ((host_pred(f, lsf) >= k) != (b >= hi)) {
                set_error("predict_scale_level: host logf is not monotone around level %d", k);
                return PLP_ERR_INVALID;
            }
        }
        memcpy(&thr[k], &hi, 4);
    }
    return PLP_OK;
}

static plp_status fill_params(FuseParams &P, const plp_camera *cam, const plp_grid *grid, const float *scale_factors,
                              const float *inv_level_sigma_sq, int num_levels, float log_scale_factor, float margin,
                              int mode) {
    memset(&P, 0, sizeof(P));
    P.cam = *cam;
    if (grid) P.grid = *grid;
    PLP_TRY(build_level_thresholds(log_scale_factor, num_levels, P.level_thr));
    for (int l = 0; l < num_levels; ++l) {
        P.scale_factors[l] = scale_factors[l];
        P.inv_sigma_sq[l] = inv_level_sigma_sq[l];
    }
    P.num_levels = num_levels;
    P.margin = margin;
    P.mode = mode;
    return PLP_OK;
}

struct LmOffsets {
    size_t pos, normal, min_d, max_d, max_raw, desc, valid;
};

static void pack_landmarks(Packer &pk, const plp_fuse_landmarks *lms, int doubles_per_lm, LmOffsets &o) {
    const size_t m = (size_t)lms->m;
    o.pos = pk.add(lms->pos_w, m * doubles_per_lm * 8);
    o.normal = pk.add(lms->obs_mean_normal, m * 3 * 8);
    o.min_d = pk.add(lms->min_valid_dist, m * 4);
    o.max_d = pk.add(lms->max_valid_dist, m * 4);
    o.max_raw = pk.add(lms->max_valid_dist_raw, m * 4);
    o.desc = pk.add(lms->desc, m * 32);
    o.valid = pk.add(lms->valid, m);
}

static FuseLandmarks bind_landmarks(uint8_t *d, const LmOffsets &o, int m) {
    FuseLandmarks L;
    L.m = m;
    L.pos_w = Packer::at<double>(d, o.pos);
    L.normal = Packer::at<double>(d, o.normal);
    L.min_d = Packer::at<float>(d, o.min_d);
    L.max_d = Packer::at<float>(d, o.max_d);
    L.max_raw = Packer::at<float>(d, o.max_raw);
    L.desc = Packer::at<uint8_t>(d, o.desc);
    L.valid = Packer::at<uint8_t>(d, o.valid);
    return L;
}

// landmarks per CTA so that (chunks x targets) fills the GPU about twice
static int chunk_size(const plp_ctx *ctx, int m, int num_targets) {
    const int want = std::max(1, (2 * std::max(ctx->sm_count, 1) + num_targets - 1) / num_targets);
    return std::max(32, div_up(m, want));
}

}  // namespace

}  // namespace plp

using namespace plp;

extern "C" {

plp_status plp_fuse_level_thresholds(float log_scale_factor, int num_levels, float *thr_out) {
    PLP_REQUIRE(thr_out, "null pointer");
    float thr[kMaxLevels];
    PLP_TRY(build_level_thresholds(log_scale_factor, num_levels, thr));
    thr_out[0] = 0.0f;
    for (int k = 1; k < num_levels; ++k) thr_out[k] = thr[k];
    return PLP_OK;
}

plp_status plp_fuse_search_points(plp_ctx *ctx, const plp_fuse_target_points *targets, int num_targets,
                                  const plp_grid *grid, const plp_camera *cam, const float *scale_factors,
                                  const float *inv_level_sigma_sq, int num_levels, float log_scale_factor,
                                  const plp_fuse_landmarks *lms, float margin, int mode, int32_t *best_idx_out,
                                  uint16_t *best_dist_out) {
    PLP_REQUIRE(ctx && grid && cam && scale_factors && inv_level_sigma_sq && lms && best_idx_out, "null pointer");
    PLP_REQUIRE(num_targets >= 0 && lms->m >= 0, "sizes");
    PLP_REQUIRE(mode == PLP_FUSE_DETECT || mode == PLP_FUSE_REPLACE, "mode");
    if (num_targets == 0 || lms->m == 0) return PLP_OK;
    PLP_REQUIRE(targets, "targets");
    PLP_REQUIRE(lms->pos_w && lms->obs_mean_normal && lms->min_valid_dist && lms->max_valid_dist &&
                    lms->max_valid_dist_raw && lms->desc,
                "landmark arrays");
    PLP_REQUIRE(grid->num_cols >= 1 && grid->num_rows >= 1 && (long long)grid->num_cols * grid->num_rows <= 16384,
                "grid size");
    const int m = lms->m;
    int max_n = 0;
    for (int t = 0; t < num_targets; ++t) {
        const plp_frame_points &f = targets[t].pts;
        PLP_REQUIRE(f.n >= 0, "target size");
        PLP_REQUIRE(f.n == 0 || (f.x && f.y && f.octave && f.desc), "target arrays");
        max_n = std::max(max_n, f.n);
    }
    if (max_n > kFuseMaxPoints) {
        set_error("fuse: %d keypoints exceed the per-keyframe capacity %d", max_n, kFuseMaxPoints);
        return PLP_ERR_CAPACITY;
    }
    FuseParams P;
    PLP_TRY(fill_params(P, cam, grid, scale_factors, inv_level_sigma_sq, num_levels, log_scale_factor, margin, mode));
    PLP_CUDA_TRY(cudaSetDevice(ctx->device));
    Packer pk;
    LmOffsets lo;
    pack_landmarks(pk, lms, 3, lo);
    struct TOff {
        size_t x, y, xr, oct, desc, skip;
    };
    std::vector<TOff> toff(num_targets);
    for (int t = 0; t < num_targets; ++t) {
        const plp_frame_points &f = targets[t].pts;
        const size_t n = (size_t)f.n;
        toff[t].x = pk.add(n ? f.x : nullptr, n * 4);
        toff[t].y = pk.add(n ? f.y : nullptr, n * 4);
        toff[t].xr = pk.add(n ? f.x_right : nullptr, n * 4);
        toff[t].oct = pk.add(n ? f.octave : nullptr, n * 4);
        toff[t].desc = pk.add(n ? f.desc : nullptr, n * 32);
        toff[t].skip = pk.add(targets[t].skip, (size_t)m);
    }
    std::vector<FusePointTarget> dev_targets(num_targets);
    const size_t o_targets = pk.add(dev_targets.data(), sizeof(FusePointTarget) * (size_t)num_targets);
    const size_t o_idx = pk.reserve((size_t)num_targets * m * 4), o_dist = pk.reserve((size_t)num_targets * m * 2);
    // device addresses are known once the scratch buffer is sized: size it first, then fill the target table in place
    void *dscratch = nullptr;
    PLP_TRY(ctx_scratch(ctx, 0, pk.total ? pk.total : 256, &dscratch));
    uint8_t *d = (uint8_t *)dscratch;
    for (int t = 0; t < num_targets; ++t) {
        FusePointTarget &T = dev_targets[t];
        memset(&T, 0, sizeof(T));
        T.n = targets[t].pts.n;
        T.x = Packer::at<float>(d, toff[t].x);
        T.y = Packer::at<float>(d, toff[t].y);
        T.xr = Packer::at<float>(d, toff[t].xr);
        T.octave = Packer::at<int32_t>(d, toff[t].oct);
        T.desc = Packer::at<uint8_t>(d, toff[t].desc);
        T.skip = Packer::at<uint8_t>(d, toff[t].skip);
        memcpy(T.R, targets[t].rot_cw, sizeof(T.R));
        memcpy(T.t, targets[t].trans_cw, sizeof(T.t));
        memcpy(T.c, targets[t].cam_center, sizeof(T.c));
    }
    uint8_t *d2;
    PLP_TRY(pk.upload(ctx, 0, &d2));
    if (d2 != d) {
        set_error("fuse: scratch buffer moved between sizing and upload");
        return PLP_ERR_CUDA;
    }
    const FuseLandmarks L = bind_landmarks(d, lo, m);
    const int cap = max_n < 64 ? 64 : ((max_n + 63) / 64) * 64;
    const size_t smem = fuse_point_smem_bytes(cap, grid->num_cols * grid->num_rows);
    PLP_CUDA_TRY(cudaFuncSetAttribute(fuse_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int chunk = chunk_size(ctx, m, num_targets);
    dim3 g(div_up(m, chunk), num_targets);
    PLP_LAUNCH(ctx, fuse_points_kernel, g, kThreads, smem, Packer::at<FusePointTarget>(d, o_targets), L, P, cap, chunk,
               Packer::at<int32_t>(d, o_idx), Packer::at<uint16_t>(d, o_dist));
    PLP_CHECK_LAUNCH();
    PLP_CUDA_TRY(cudaMemcpyAsync(best_idx_out, d + o_idx, (size_t)num_targets * m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (best_dist_out)
        PLP_CUDA_TRY(cudaMemcpyAsync(best_dist_out, d + o_dist, (size_t)num_targets * m * 2, cudaMemcpyDeviceToHost,
                                     ctx->stream));
    PLP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PLP_OK;
}

plp_status plp_fuse_search_lines(plp_ctx *ctx, const plp_fuse_target_lines *targets, int num_targets,
                                 const plp_camera *cam, const float *scale_factors_lsd,
                                 const float *inv_level_sigma_sq_lsd, int num_levels_lsd, float log_scale_factor_lsd,
                                 const plp_fuse_landmarks *lms, float margin, int32_t *best_idx_out,
                                 uint16_t *best_dist_out) {
    PLP_REQUIRE(ctx && cam && scale_factors_lsd && inv_level_sigma_sq_lsd && lms && best_idx_out, "null pointer");
    PLP_REQUIRE(num_targets >= 0 && lms->m >= 0, "sizes");
    if (num_targets == 0 || lms->m == 0) return PLP_OK;
    PLP_REQUIRE(targets, "targets");
    PLP_REQUIRE(lms->pos_w && lms->min_valid_dist && lms->max_valid_dist && lms->max_valid_dist_raw && lms->desc,
                "landmark arrays");
    const int m = lms->m;
    int max_n = 0;
    for (int t = 0; t < num_targets; ++t) {
        const plp_frame_lines &f = targets[t].lines;
        PLP_REQUIRE(f.n >= 0, "target size");
        PLP_REQUIRE(f.n == 0 || (f.sx && f.sy && f.ex && f.ey && f.octave && f.desc), "target arrays");
        max_n = std::max(max_n, f.n);
    }
    if (max_n > kFuseMaxLines) {
        set_error("fuse: %d keylines exceed the per-keyframe capacity %d", max_n, kFuseMaxLines);
        return PLP_ERR_CAPACITY;
    }
    FuseParams P;
    PLP_TRY(fill_params(P, cam, nullptr, scale_factors_lsd, inv_level_sigma_sq_lsd, num_levels_lsd, log_scale_factor_lsd,
                        margin, PLP_FUSE_REPLACE));
    PLP_CUDA_TRY(cudaSetDevice(ctx->device));
    Packer pk;
    LmOffsets lo;
    plp_fuse_landmarks lms_no_normal = *lms;
    lms_no_normal.obs_mean_normal = nullptr;
    pack_landmarks(pk, &lms_no_normal, 6, lo);
    struct TOff {
        size_t sx, sy, ex, ey, oct, desc, skip;
    };
    std::vector<TOff> toff(num_targets);
    for (int t = 0; t < num_targets; ++t) {
        const plp_frame_lines &f = targets[t].lines;
        const size_t n = (size_t)f.n;
        toff[t].sx = pk.add(n ? f.sx : nullptr, n * 4);
        toff[t].sy = pk.add(n ? f.sy : nullptr, n * 4);
        toff[t].ex = pk.add(n ? f.ex : nullptr, n * 4);
        toff[t].ey = pk.add(n ? f.ey : nullptr, n * 4);
        toff[t].oct = pk.add(n ? f.octave : nullptr, n * 4);
        toff[t].desc = pk.add(n ? f.desc : nullptr, n * 32);
        toff[t].skip = pk.add(targets[t].skip, (size_t)m);
    }
    std::vector<FuseLineTarget> dev_targets(num_targets);
    const size_t o_targets = pk.add(dev_targets.data(), sizeof(FuseLineTarget) * (size_t)num_targets);
    const size_t o_idx = pk.reserve((size_t)num_targets * m * 4), o_dist = pk.reserve((size_t)num_targets * m * 2);
    void *dscratch = nullptr;
    PLP_TRY(ctx_scratch(ctx, 0, pk.total ? pk.total : 256, &dscratch));
    uint8_t *d = (uint8_t *)dscratch;
    for (int t = 0; t < num_targets; ++t) {
        FuseLineTarget &T = dev_targets[t];
        memset(&T, 0, sizeof(T));
        T.n = targets[t].lines.n;
        T.sx = Packer::at<float>(d, toff[t].sx);
        T.sy = Packer::at<float>(d, toff[t].sy);
        T.ex = Packer::at<float>(d, toff[t].ex);
        T.ey = Packer::at<float>(d, toff[t].ey);
        T.octave = Packer::at<int32_t>(d, toff[t].oct);
        T.desc = Packer::at<uint8_t>(d, toff[t].desc);
        T.skip = Packer::at<uint8_t>(d, toff[t].skip);
        memcpy(T.R, targets[t].rot_cw, sizeof(T.R));
        memcpy(T.t, targets[t].trans_cw, sizeof(T.t));
        memcpy(T.c, targets[t].cam_center, sizeof(T.c));
    }
    uint8_t *d2;
    PLP_TRY(pk.upload(ctx, 0, &d2));
    if (d2 != d) {
        set_error("fuse: scratch buffer moved between sizing and upload");
        return PLP_ERR_CUDA;
    }
    const FuseLandmarks L = bind_landmarks(d, lo, m);
    const int cap = max_n < 64 ? 64 : ((max_n + 63) / 64) * 64;
    const size_t smem = (size_t)cap * (32 + 5 * 4);
    PLP_CUDA_TRY(cudaFuncSetAttribute(fuse_lines_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int chunk = chunk_size(ctx, m, num_targets);
    dim3 g(div_up(m, chunk), num_targets);
    PLP_LAUNCH(ctx, fuse_lines_kernel, g, kThreads, smem, Packer::at<FuseLineTarget>(d, o_targets), L, P, cap, chunk,
               Packer::at<int32_t>(d, o_idx), Packer::at<uint16_t>(d, o_dist));
    PLP_CHECK_LAUNCH();
    PLP_CUDA_TRY(cudaMemcpyAsync(best_idx_out, d + o_idx, (size_t)num_targets * m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (best_dist_out)
        PLP_CUDA_TRY(cudaMemcpyAsync(best_dist_out, d + o_dist, (size_t)num_targets * m * 2, cudaMemcpyDeviceToHost,
                                     ctx->stream));
    PLP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PLP_OK;
}

}  // extern "C"
