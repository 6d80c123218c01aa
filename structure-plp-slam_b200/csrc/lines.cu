// lines.cu -- LSD line segments + LBD descriptors for batches of frames, sm_100a.
//
// Replaces LineFeatureTracker::extract_LSD_LBD (feature/line_extractor.cc:88-160) and what it calls:
//   LSDDetectorC::detectImpl            feature/line_descriptor/LSDDetector_custom.cpp:225-320
//   cv::LineSegmentDetector::detect     third party (OpenCV imgproc lsd.cpp); LSD of Grompone von Gioi et al., IPOL 2012,
//                                       refine = LSD_REFINE_STD, options of line_extractor.cc:113-122
//   BinaryDescriptor::compute           feature/line_descriptor/binary_descriptor_custom.cpp:518-679, 1018-1364
//
// LSD is a sequential greedy algorithm per image (a seed grows a region over a shared `used` map, the running region
// angle decides every next pixel), so the parallel axes are (1) the frames of a batch -- one warp owns one frame's
// region growing -- and (2) inside the warp: the 3 x 8 neighbours of three queue entries are fetched by 27 lanes at once
// and tested against the current region angle in parallel; only the accepted pixels are applied in sequence (the first
// aligned lane is what the scalar loop would have taken; later lanes are re-tested against the updated angle), which
// reproduces the scalar visiting order exactly.  Rectangle fitting / refinement are warp reductions with a fixed
// summation order.  Everything around it (11x11 blur + 1/2 down-scale, gradient + level-line angle, the 1024-bin stable
// seed ordering, 5x5 blur + Sobel, the 63-row LBD band sums) is ordinary data-parallel work.
//
// Determinism rules (the parity checker restates the same rules and is itself pinned bit-exactly against cv2 4.13):
// seeds ordered by (bin desc, raster asc); cos/sin/atan2 from detmath.h; double sums as 32 strided partials + xor tree;
// order-preserving compaction in reduce_region_radius.  Compiled with -fmad=false.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "detmath.h"

using namespace plp;

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr double kDegToRads = 0.017453292519943295769236907684;
constexpr double kPi = 3.14159265358979323846;
constexpr double k3_2Pi = 4.71238898038;  // literals of lsd.cpp
constexpr double k2Pi = 6.28318530718;
constexpr int kBins = 1024;
constexpr int kSortWarps = 32;
constexpr int kBands = 9, kBandWidth = 7, kLspHeight = kBands * kBandWidth;

struct LineDev {
    int w, h;          // full resolution
    int sw, sh, npx;   // half resolution LSD works on
    int seg_cap, kl_cap;
    int min_reg_size;
    double rho, prec, p, density_th, min_length;
    // per batch buffers (frame-major)
    const uint8_t *img;
    size_t img_step, img_frame_stride;
    uint8_t *scaled;     // npx
    const float4 *cstab; // (2*510+1)^2 x {deg, cos, sin} by (gx, gy), shared by all frames
    int kthr;            // level-line angle defined  <=>  gx^2+gy^2 > kthr  (norm > rho)
    uint32_t *order;     // npx packed (y<<16|x) seeds, bin desc / raster asc
    int *nseeds;
    uint32_t *reg_xy;    // npx: region entries beyond the shared-memory window
    int reg_cap_small;   // region window (entries) of lsd_grow_kernel<false>
    int direct_trig;     // bit 0: lsd_grow_mw_kernel, bit 1: lsd_grow_kernel compute the neighbour's {deg, cos, sin} directly
    unsigned long long *mw_stat;  // per frame {rounds, seeds run, seeds redone} of lsd_grow_mw_kernel (may be null)
    float4 *segs;        // seg_cap
    int *nseg;
    short2 *grad;        // w*h Sobel (dx, dy) of the 5x5-blurred frame
    float *lbd_float;    // kl_cap x 72
    int *status;
    float gauss_l[kBandWidth * 3];
    float gauss_g[kLspHeight];
};

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        if (p >= len) p = 2 * (len - 1) - p;
    }
    return p;
}

// cv::fastAtan2 (degrees), f32 without FMA
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale, p5 = 0.1555786518463281f * scale,
                p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    // branch-free form of `if (ax >= ay) c = ay / (ax + eps) else c = ax / (ay + eps)`: one division for all lanes
    const bool steep = !(ax >= ay);
    const float mn = steep ? ax : ay, mx = steep ? ay : ax;
    const float c = mn / (mx + 2.220446049250313e-16f);
    const float c2 = c * c;
    float a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    if (steep) a = 90.f - a;
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ------------------------------------------------------------------------------------------------------------------
// K1: GaussianBlur(11x11, sigma 1.2) in OpenCV's Q8 fixed point (taps 0 0 4 21 60 86 60 21 4 0 0) followed by
//     resize(0.5, INTER_LINEAR_EXACT) == rounded mean of each 2x2 block.  One CTA = 64 x 16 source pixels.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kScTw = 64, kScTh = 16;
__global__ void __launch_bounds__(256) lsd_scale_kernel(LineDev D) {
    __shared__ uint8_t s_src[kScTh + 6][kScTw + 8];
    __shared__ uint16_t s_h[kScTh + 6][kScTw];
    __shared__ uint8_t s_b[kScTh][kScTw];
    const int b = blockIdx.y;
    const int tiles_x = (D.w + kScTw - 1) / kScTw;
    const int tx = (blockIdx.x % tiles_x) * kScTw, ty = (blockIdx.x / tiles_x) * kScTh;
    const uint8_t *img = D.img + (size_t)b * D.img_frame_stride;
    for (int i = threadIdx.x; i < (kScTh + 6) * (kScTw + 6); i += blockDim.x) {
        const int r = i / (kScTw + 6), c = i - r * (kScTw + 6);
        const int y = reflect101(ty + r - 3, D.h), x = reflect101(tx + c - 3, D.w);
        s_src[r][c] = img[(size_t)y * D.img_step + x];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (kScTh + 6) * kScTw; i += blockDim.x) {
        const int r = i / kScTw, c = i - r * kScTw;
        const uint8_t *s = &s_src[r][c];
        s_h[r][c] = (uint16_t)(4 * (s[0] + s[6]) + 21 * (s[1] + s[5]) + 60 * (s[2] + s[4]) + 86 * s[3]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kScTh * kScTw; i += blockDim.x) {
        const int r = i / kScTw, c = i - r * kScTw;
        const uint32_t v = 4u * (s_h[r][c] + s_h[r + 6][c]) + 21u * (s_h[r + 1][c] + s_h[r + 5][c]) +
                           60u * (s_h[r + 2][c] + s_h[r + 4][c]) + 86u * s_h[r + 3][c];
        s_b[r][c] = (uint8_t)((v + 32768u) >> 16);
    }
    __syncthreads();
    uint8_t *out = D.scaled + (size_t)b * D.npx;
    for (int i = threadIdx.x; i < (kScTh / 2) * (kScTw / 2); i += blockDim.x) {
        const int r = i / (kScTw / 2), c = i - r * (kScTw / 2);
        const int ox = tx / 2 + c, oy = ty / 2 + r;
        if (ox >= D.sw || oy >= D.sh) continue;
        // clamp like the oracle for odd sizes (min(2x+1, w-1)); inside the tile the clamped pixel is the same column/row
        const int x0 = 2 * c, x1 = (tx + 2 * c + 1 < D.w) ? 2 * c + 1 : 2 * c;
        const int y0 = 2 * r, y1 = (ty + 2 * r + 1 < D.h) ? 2 * r + 1 : 2 * r;
        const int s = s_b[y0][x0] + s_b[y0][x1] + s_b[y1][x0] + s_b[y1][x1];
        out[(size_t)oy * D.sw + ox] = (uint8_t)((s + 2) >> 2);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// ll_angle helpers.  The level-line field is never materialised: every consumer recomputes the 2x2 gradient from the
// half-resolution image (4 byte reads), which is what lets one frame's working set (image + `used` bitmap) live in
// shared memory.  (gx, gy) in [-510, 510]^2 determines the angle, hence cos/sin come from a table indexed by (gx, gy)
// that is built once per handle and shared by every frame (8.3 MB, L2 resident).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGRange = 510, kGDim = 2 * kGRange + 1;

__device__ __forceinline__ void grad_at(const uint8_t *img, int sw, int idx, int &gx, int &gy) {
    const int a = img[idx], bq = img[idx + 1], c = img[idx + sw], d = img[idx + sw + 1];
    const int DA = d - a, BC = bq - c;
    gx = DA + BC;
    gy = DA - BC;
}

// table entry: level-line angle in degrees (cv::fastAtan2(gx, -gy)) and cos / sin of float(angle) as lsd.cpp
// accumulates them: `sumdx += cos(float(angle))`
__global__ void __launch_bounds__(256) lsd_cs_table_kernel(float4 *tab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kGDim * kGDim) return;
    const int gy = i / kGDim - kGRange, gx = i - (gy + kGRange) * kGDim - kGRange;
    const float deg = fast_atan2_deg((float)gx, (float)-gy);
    const double a = (double)deg * kDegToRads;
    const float af = (float)a;
    tab[i] = make_float4(deg, (float)det_cos((double)af), (float)det_sin((double)af), 0.f);
}

// ------------------------------------------------------------------------------------------------------------------
// K3: pseudo-ordering of the seeds: bin = int(modgrad * 1023 / max_grad), descending bins, raster order inside a bin.
//     One CTA (32 warps) per frame: frame maximum, then each warp owns a contiguous raster range -> per-warp
//     histograms, a scan over (bin desc, warp asc), and a stable scatter with __match_any ranks.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int seed_k(const uint8_t *img, int sw, int sh, int i, int kthr) {
    const int y = i / sw, x = i - y * sw;
    if (x >= sw - 1 || y >= sh - 1) return -1;
    int gx, gy;
    grad_at(img, sw, i, gx, gy);
    const int k = gx * gx + gy * gy;
    return k > kthr ? k : -1;  // norm <= rho  <=>  k <= kthr
}

__global__ void __launch_bounds__(kSortWarps * 32) lsd_sort_kernel(LineDev D) {
    extern __shared__ uint32_t s_dyn[];
    uint32_t *hist = s_dyn;                        // [kSortWarps][kBins]: counts, then running start offsets
    uint32_t *base = s_dyn + kSortWarps * kBins;   // [kBins] first output slot of each bin
    __shared__ uint32_t s_scan[kBins];
    __shared__ uint32_t s_warp_tot[32];
    __shared__ int s_kmax[32];
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint8_t *img = D.scaled + (size_t)b * D.npx;
    // frame maximum of the squared norm over the defined pixels
    int kloc = 0;
    for (int i = threadIdx.x; i < D.npx; i += blockDim.x) kloc = max(kloc, seed_k(img, D.sw, D.sh, i, D.kthr));
    for (int off = 16; off >= 1; off >>= 1) kloc = max(kloc, __shfl_xor_sync(kFull, kloc, off));
    if (lane == 0) s_kmax[wid] = kloc;
    for (int i = threadIdx.x; i < kSortWarps * kBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    int kmax = 0;
    for (int wv = 0; wv < kSortWarps; ++wv) kmax = max(kmax, s_kmax[wv]);
    const double max_grad = kmax > 0 ? sqrt((double)kmax / 4.0) : -1.0;
    const double bin_coef = (max_grad > 0) ? (double)(kBins - 1) / max_grad : 0.0;
    const int per_warp = ((D.npx + kSortWarps * 32 - 1) / (kSortWarps * 32)) * 32;
    const int beg = wid * per_warp, end = min(beg + per_warp, D.npx);
    uint32_t *myhist = hist + wid * kBins;
    // pass 1: per-warp histogram (lanes of one warp may hit the same bin: one leader per bin adds the group size)
    for (int i0 = beg; i0 < end; i0 += 32) {
        const int i = i0 + lane;
        int bin = -1;
        if (i < end) {
            const int k = seed_k(img, D.sw, D.sh, i, D.kthr);
            if (k >= 0) bin = (int)(sqrt((double)k / 4.0) * bin_coef);
        }
        const unsigned act = __ballot_sync(kFull, bin >= 0);
        if (bin >= 0) {
            const unsigned peers = __match_any_sync(act, bin);
            if (lane == __ffs(peers) - 1) myhist[bin] += (uint32_t)__popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    // pass 2: thread t owns bin t: per-warp exclusive starts inside the bin, bin total; then an exclusive scan over the
    // bins in DESCENDING order (rank = 1023 - bin)
    {
        const int bin = threadIdx.x;
        uint32_t run = 0;
        for (int wv = 0; wv < kSortWarps; ++wv) {
            const uint32_t c = hist[wv * kBins + bin];
            hist[wv * kBins + bin] = run;
            run += c;
        }
        s_scan[kBins - 1 - bin] = run;
    }
    __syncthreads();
    {
        const uint32_t v = s_scan[threadIdx.x];  // total of rank threadIdx.x
        uint32_t incl = v;
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t nb = __shfl_up_sync(kFull, incl, off);
            if (lane >= off) incl += nb;
        }
        if (lane == 31) s_warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const uint32_t t = s_warp_tot[lane];
            uint32_t ti = t;
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t nb = __shfl_up_sync(kFull, ti, off);
                if (lane >= off) ti += nb;
            }
            s_warp_tot[lane] = ti - t;
            if (lane == 31) D.nseeds[b] = (int)ti;
        }
        __syncthreads();
        base[kBins - 1 - threadIdx.x] = incl - v + s_warp_tot[wid];
    }
    __syncthreads();
    // pass 3: stable scatter
    uint32_t *order = D.order + (size_t)b * D.npx;
    for (int i0 = beg; i0 < end; i0 += 32) {
        const int i = i0 + lane;
        int bin = -1;
        if (i < end) {
            const int k = seed_k(img, D.sw, D.sh, i, D.kthr);
            if (k >= 0) bin = (int)(sqrt((double)k / 4.0) * bin_coef);
        }
        const unsigned act = __ballot_sync(kFull, bin >= 0);
        if (bin >= 0) {
            const unsigned peers = __match_any_sync(act, bin);
            const int rank = __popc(peers & ((1u << lane) - 1));
            const uint32_t start = myhist[bin];
            const int y = i / D.sw, x = i - y * D.sw;
            order[base[bin] + start + rank] = ((uint32_t)y << 16) | (uint32_t)x;
            __syncwarp(act);
            if (lane == __ffs(peers) - 1) myhist[bin] = start + (uint32_t)__popc(peers);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K4: region growing + rectangle + refinement: one warp per frame; half-resolution image, `used` bitmap and the region
//     list (= BFS queue) in shared memory.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kRegCap = 6144;       // region entries kept in shared memory; longer regions spill to global memory
constexpr int kRegCapSmall = 2048;  // ... in the many-frames-per-SM variant

struct Rect {
    double x1, y1, x2, y2, width;
};

// kMw = false: one warp per frame, marks go straight into the frame's `used` bitmap.
// kMw = true (lsd_grow_mw_kernel): several warps of a CTA work on seeds of the SAME frame speculatively: `used` is the
// committed bitmap (read only while the warps grow), the warp's own marks live in its private bitmap `mark`, and the
// bounding box of every pixel the warp ever accepted is tracked for the conflict test.
template <bool kMw>
struct GrowT {  // per-warp state
    int sw, sh, kthr;
    double density_th;
    const uint8_t *img;    // shared: half-resolution image
    uint32_t *used;        // shared: bitmap (kMw: committed marks of the frame)
    uint32_t *mark;        // shared: where this warp sets / clears marks (= used unless kMw)
    uint32_t *reg;         // shared: first reg_cap region entries (packed y<<16|x)
    uint32_t *reg_ovf;     // global: all entries beyond reg_cap (indexed by absolute position)
    const float4 *tab;     // global: {deg, cos, sin} by (gx, gy)
    int lane, reg_cap;
    bool direct;           // compute {deg, cos, sin} of a neighbour instead of reading the table
    mutable int bx0, by0, bx1, by1;  // kMw: per-lane bounding box of the pixels this lane accepted (reduced by the caller)
    // out-of-order kernel only (claim == nullptr otherwise): `claim` = union of the private marks of every context in flight;
    // a region that is about to accept a pixel claimed by an EARLIER ticket stops at once (`aborted`) and is decided at the head
    uint32_t *claim;
    const uint32_t *priv_base;   // the contexts' private bitmaps, `ctx_words` words apart
    const int *ctx_ticket;       // ticket each context is working on (INT_MAX: idle)
    int nctx, self, my_ticket, ctx_words;
    mutable bool aborted;
#ifdef PLP_LSD_PROF
    long long *pc;         // [0] iterations [1] rounds [2] cycles load phase [3] cycles resolve phase [4] on-demand loads
#endif
    __device__ __forceinline__ uint32_t get(int e) const { return e < reg_cap ? reg[e] : reg_ovf[e]; }
    __device__ __forceinline__ void put(int e, uint32_t v) const {
        if (e < reg_cap) reg[e] = v;
        else reg_ovf[e] = v;
    }
    __device__ __forceinline__ bool is_used(int idx) const {
        uint32_t w = used[idx >> 5];
        if (kMw) w |= mark[idx >> 5];
        return (w >> (idx & 31)) & 1u;
    }
    __device__ __forceinline__ bool claimed_by_earlier(int idx) const {
        if (!kMw || !claim || !((claim[idx >> 5] >> (idx & 31)) & 1u)) return false;
        bool earlier = false;
        for (int w = 0; w < nctx; ++w)
            if (w != self && ((priv_base[(size_t)w * ctx_words + (idx >> 5)] >> (idx & 31)) & 1u))
                earlier = earlier || *reinterpret_cast<const volatile int *>(&ctx_ticket[w]) < my_ticket;
        return earlier;
    }
    __device__ __forceinline__ void unclaim(int idx) const {
        if (kMw && claim) atomicAnd(&claim[idx >> 5], ~(1u << (idx & 31)));
    }
    __device__ __forceinline__ void accept(int idx, uint32_t xy) const {  // one lane: mark a pixel of the region
        mark[idx >> 5] |= 1u << (idx & 31);
        if (kMw && claim) atomicOr(&claim[idx >> 5], 1u << (idx & 31));
        if (kMw) {
            const int x = (int)(xy & 0xffff), y = (int)(xy >> 16);
            bx0 = min(bx0, x);
            bx1 = max(bx1, x);
            by0 = min(by0, y);
            by1 = max(by1, y);
        }
    }
};
using Grow = GrowT<false>;

__device__ __forceinline__ bool is_aligned(double a, double theta, double prec) {
    double n_theta = theta - a;
    if (n_theta < 0) n_theta = -n_theta;
    if (n_theta > k3_2Pi) {
        n_theta -= k2Pi;
        if (n_theta < 0) n_theta = -n_theta;
    }
    return n_theta <= prec;
}

__device__ __forceinline__ double warp_sum_tree(double p) {
    for (int off = 16; off >= 1; off >>= 1) p = p + __shfl_xor_sync(kFull, p, off);
    return p;
}
__device__ __forceinline__ double warp_max(double p) {
    for (int off = 16; off >= 1; off >>= 1) p = fmax(p, __shfl_xor_sync(kFull, p, off));
    return p;
}
__device__ __forceinline__ double warp_min(double p) {
    for (int off = 16; off >= 1; off >>= 1) p = fmin(p, __shfl_xor_sync(kFull, p, off));
    return p;
}

// immutable data of one neighbour pixel (does not depend on the `used` map)
struct Nb {
    int nidx;     // -1: outside / no gradient defined
    uint32_t xy;
    float4 t;     // {deg, cos, sin}
};

// neighbour jj (0..7, centre skipped) of queue entry e
template <bool kMw>
__device__ __forceinline__ Nb load_nb(const GrowT<kMw> &G, int e, int ddx, int ddy) {
    Nb r;
    r.nidx = -1;
    r.xy = 0;
    r.t = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t pxy = G.get(e);
    const int nx = (int)(pxy & 0xffff) + ddx, ny = (int)(pxy >> 16) + ddy;
    // the last row / column carry no gradient (NOTDEF)
    if (nx >= 0 && nx < G.sw - 1 && ny >= 0 && ny < G.sh - 1) {
        const int idx = ny * G.sw + nx;
        int gx, gy;
        grad_at(G.img, G.sw, idx, gx, gy);
        if (gx * gx + gy * gy > G.kthr) {
            r.nidx = idx;
            r.xy = ((uint32_t)ny << 16) | (uint32_t)nx;
            if (G.direct) {  // latency mode: ~250 dependent cycles of arithmetic instead of a table entry from L2
                const float deg = fast_atan2_deg((float)gx, (float)-gy);
                const double af = (double)(float)((double)deg * kDegToRads);
                r.t = make_float4(deg, (float)det_cos(af), (float)det_sin(af), 0.f);
            } else {
                r.t = G.tab[(gy + kGRange) * kGDim + gx + kGRange];
            }
        }
    }
    return r;
}

// lsd.cpp region_grow.  Returns the region size; the region list lives in G.reg (+ overflow).
// 32 lanes = 4 queue entries x 8 neighbours, in the scalar visiting order (entry, then yy, then xx).  The immutable data
// of the next four entries is fetched while the current four are resolved; every candidate lane keeps the region sums
// and angle it WOULD produce if it were accepted next, so an acceptance is one shuffle away.
template <bool kMw>
__device__ int region_grow(const GrowT<kMw> &G, uint32_t seed_xy, float seed_deg, double prec, double &reg_angle_out) {
    const int sw = G.sw, lane = G.lane;
    double reg_angle = (double)seed_deg * kDegToRads;
    float sumdx = (float)det_cos(reg_angle);
    float sumdy = (float)det_sin(reg_angle);
    if (lane == 0) {
        const int sidx = (int)(seed_xy >> 16) * sw + (int)(seed_xy & 0xffff);
        G.accept(sidx, seed_xy);
        G.reg[0] = seed_xy;
    }
    __syncwarp();
    int n = 1, i = 0;
    const int g = lane >> 3, jj = lane & 7;
    const int j = jj + (jj >= 4);  // skip the centre
    const int ddx = j % 3 - 1, ddy = j / 3 - 1;
    Nb cur;
    cur.nidx = -1;
    cur.xy = 0;
    cur.t = make_float4(0.f, 0.f, 0.f, 0.f);
    int loaded = 0;  // groups of `cur` that hold valid data
    while (i < n) {
        const int take = min(4, n - i);
#ifdef PLP_LSD_PROF
        const long long tl0 = clock64();
        G.pc[0]++;
        if (loaded < take) G.pc[4]++;
#endif
        if (g >= loaded && g < take) cur = load_nb(G, i + g, ddx, ddy);  // entries that were not known one round ago
        // prefetch the entries already known for the next round
        const int nxt_avail = min(4, n - (i + take));
        Nb nxt;
        nxt.nidx = -1;
        nxt.xy = 0;
        nxt.t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < nxt_avail) nxt = load_nb(G, i + take + g, ddx, ddy);
        // resolve the current entries
#ifdef PLP_LSD_PROF
        const long long tl1 = clock64();
        G.pc[2] += tl1 - tl0;
#endif
        bool cand = (g < take) && (cur.nidx >= 0) && !G.is_used(cur.nidx);
        const bool ce = cand && G.claimed_by_earlier(cur.nidx);  // (false unless the out-of-order kernel runs)
        const double a = (double)cur.t.x * kDegToRads;
        float my_sdx = sumdx + cur.t.y, my_sdy = sumdy + cur.t.z;
        double my_theta = (double)fast_atan2_deg(my_sdy, my_sdx) * kDegToRads;
        for (;;) {
            const bool al = cand && is_aligned(a, reg_angle, prec);
            const unsigned m = __ballot_sync(kFull, al);
            if (!m) break;
#ifdef PLP_LSD_PROF
            G.pc[1]++;
#endif
            const int l = __ffs(m) - 1;
            if (kMw && __shfl_sync(kFull, ce ? 1 : 0, l)) {  // the next pixel of the sequential order belongs to an earlier region in flight
                G.aborted = true;
                reg_angle_out = reg_angle;
                return n;
            }
            sumdx = __shfl_sync(kFull, my_sdx, l);
            sumdy = __shfl_sync(kFull, my_sdy, l);
            reg_angle = __shfl_sync(kFull, my_theta, l);
            const int q = __shfl_sync(kFull, cur.nidx, l);
            if (lane == l) {
                G.accept(cur.nidx, cur.xy);
                G.put(n, cur.xy);
            }
            ++n;
            cand = cand && (lane > l) && (cur.nidx != q);
            my_sdx = sumdx + cur.t.y;
            my_sdy = sumdy + cur.t.z;
            my_theta = (double)fast_atan2_deg(my_sdy, my_sdx) * kDegToRads;
        }
        __syncwarp();
#ifdef PLP_LSD_PROF
        G.pc[3] += clock64() - tl1;
#endif
        i += take;
        cur = nxt;
        loaded = nxt_avail;
    }
    reg_angle_out = reg_angle;
    return n;
}

__device__ __forceinline__ double angle_diff_signed(double a, double b) {
    double diff = a - b;
    while (diff <= -kPi) diff += k2Pi;
    while (diff > kPi) diff -= k2Pi;
    return diff;
}
__device__ __forceinline__ double dist2(double x1, double y1, double x2, double y2) {
    return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1);
}
// modgrad of a region pixel, recomputed from the image
template <bool kMw>
__device__ __forceinline__ double px_weight(const GrowT<kMw> &G, uint32_t xy) {
    int gx, gy;
    grad_at(G.img, G.sw, (int)(xy >> 16) * G.sw + (int)(xy & 0xffff), gx, gy);
    return sqrt((double)(gx * gx + gy * gy) / 4.0);
}

// lsd.cpp region2rect + get_theta
template <bool kMw>
__device__ void region2rect(const GrowT<kMw> &G, int n, double reg_angle, double prec, Rect &R) {
    const int lane = G.lane;
    double sx = 0, sy = 0, ss = 0;
    for (int i = lane; i < n; i += 32) {
        const uint32_t xy = G.get(i);
        const double wgt = px_weight(G, xy);
        sx += (double)(int)(xy & 0xffff) * wgt;
        sy += (double)(int)(xy >> 16) * wgt;
        ss += wgt;
    }
    sx = warp_sum_tree(sx);
    sy = warp_sum_tree(sy);
    ss = warp_sum_tree(ss);
    const double x = sx / ss, y = sy / ss;
    double ixx = 0, iyy = 0, ixy = 0;
    for (int i = lane; i < n; i += 32) {
        const uint32_t xy = G.get(i);
        const double wgt = px_weight(G, xy);
        const double dx = (double)(int)(xy & 0xffff) - x, dy = (double)(int)(xy >> 16) - y;
        ixx += dy * dy * wgt;
        iyy += dx * dx * wgt;
        ixy += dx * dy * wgt;
    }
    const double Ixx = warp_sum_tree(ixx), Iyy = warp_sum_tree(iyy), Ixy = -warp_sum_tree(ixy);
    const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_deg((float)(lambda - Ixx), (float)Ixy)
                                           : (double)fast_atan2_deg((float)Ixy, (float)(lambda - Iyy));
    theta *= kDegToRads;
    if (fabs(angle_diff_signed(theta, reg_angle)) > prec) theta += kPi;
    const double dx = det_cos(theta), dy = det_sin(theta);
    double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
    for (int i = lane; i < n; i += 32) {
        const uint32_t xy = G.get(i);
        const double rdx = (double)(int)(xy & 0xffff) - x, rdy = (double)(int)(xy >> 16) - y;
        const double l = rdx * dx + rdy * dy;
        const double wv = -rdx * dy + rdy * dx;
        l_max = fmax(l_max, l);
        l_min = fmin(l_min, l);
        w_max = fmax(w_max, wv);
        w_min = fmin(w_min, wv);
    }
    l_max = warp_max(l_max);
    l_min = warp_min(l_min);
    w_max = warp_max(w_max);
    w_min = warp_min(w_min);
    R.x1 = x + l_min * dx;
    R.y1 = y + l_min * dy;
    R.x2 = x + l_max * dx;
    R.y2 = y + l_max * dy;
    R.width = w_max - w_min;
    if (R.width < 1.0) R.width = 1.0;
}

__device__ __forceinline__ double rect_density(int n, const Rect &R) {
    return (double)n / (sqrt(dist2(R.x1, R.y1, R.x2, R.y2)) * R.width);
}

// lsd.cpp refine + reduce_region_radius; n is updated; returns false when the region is rejected
template <bool kMw>
__device__ bool refine(const GrowT<kMw> &G, int &n, float seed_deg, double reg_angle, double prec, Rect &R) {
    const int lane = G.lane, sw = G.sw;
    double density = rect_density(n, R);
    if (density >= G.density_th) return true;
    const uint32_t seed_xy = G.reg[0];
    const double xc = (double)(int)(seed_xy & 0xffff), yc = (double)(int)(seed_xy >> 16);
    const double ang_c = (double)seed_deg * kDegToRads;
    double sum = 0, s_sum = 0;
    int cnt = 0;
    for (int i = lane; i < n; i += 32) {
        const uint32_t xy = G.get(i);
        const int px = xy & 0xffff, py = xy >> 16;
        const int pidx = py * sw + px;
        atomicAnd(&G.mark[pidx >> 5], ~(1u << (pidx & 31)));  // NOTUSED again
        G.unclaim(pidx);
        if (sqrt(dist2(xc, yc, (double)px, (double)py)) < R.width) {
            int gx, gy;
            grad_at(G.img, sw, pidx, gx, gy);
            const double ang = (double)fast_atan2_deg((float)gx, (float)-gy) * kDegToRads;
            const double d = angle_diff_signed(ang, ang_c);
            sum += d;
            s_sum += d * d;
            ++cnt;
        }
    }
    sum = warp_sum_tree(sum);
    s_sum = warp_sum_tree(s_sum);
    for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(kFull, cnt, off);
    const double mean_angle = sum / (double)cnt;
    const double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)cnt + mean_angle * mean_angle);
    __syncwarp();
    n = region_grow(G, seed_xy, seed_deg, tau, reg_angle);
    if (kMw && G.aborted) return false;
    if (n < 2) return false;
    region2rect(G, n, reg_angle, prec, R);
    density = rect_density(n, R);
    if (density >= G.density_th) return true;
    // reduce_region_radius
    const double r1 = dist2(xc, yc, R.x1, R.y1), r2 = dist2(xc, yc, R.x2, R.y2);
    double rad_sq = r1 > r2 ? r1 : r2;
    while (density < G.density_th) {
        rad_sq *= 0.75 * 0.75;
        int o = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            uint32_t xy = 0;
            bool keep = false;
            if (i < n) {
                xy = G.get(i);
                const int px = xy & 0xffff, py = xy >> 16;
                keep = !(dist2(xc, yc, (double)px, (double)py) > rad_sq);
                if (!keep) {
                    const int pidx = py * sw + px;
                    atomicAnd(&G.mark[pidx >> 5], ~(1u << (pidx & 31)));
                    G.unclaim(pidx);
                }
            }
            const unsigned km = __ballot_sync(kFull, keep);
            __syncwarp();
            if (keep) G.put(o + __popc(km & ((1u << lane) - 1)), xy);
            o += __popc(km);
            __syncwarp();
        }
        n = o;
        if (n < 2) return false;
        region2rect(G, n, reg_angle, prec, R);
        density = rect_density(n, R);
    }
    return true;
}

#ifdef PLP_LSD_PROF
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(slot, t0) prof[slot] += clock64() - (t0)
#else
#define PROF_T(var)
#define PROF_ADD(slot, t0)
#endif

// kImgSmem: the half-resolution image is staged in shared memory (lowest latency, 2 frames per SM at VGA) or read from
// global memory through L1 / L2 (34 KB of shared memory per frame -> 6 frames per SM: more frames in flight for big
// batches; the images of a batch, 77 KB each, stay L2 resident)
template <bool kImgSmem>
__global__ void __launch_bounds__(32) lsd_grow_kernel(LineDev D) {
    extern __shared__ uint4 s_grow[];
#ifdef PLP_LSD_PROF
    long long prof[6] = {0, 0, 0, 0, 0, 0};
    long long cnt_regions = 0, cnt_px = 0, cnt_refine = 0;
    __shared__ long long s_pc[8];
    for (int q = 0; q < 8; ++q) s_pc[q] = 0;
    const long long t_start = clock64();
#endif
    uint8_t *s_img = reinterpret_cast<uint8_t *>(s_grow);
    const int img_bytes = kImgSmem ? ((D.npx + 15) & ~15) : 0;
    uint32_t *s_used = reinterpret_cast<uint32_t *>(s_img + img_bytes);
    const int used_words = (D.npx + 31) >> 5;
    uint32_t *s_reg = s_used + ((used_words + 3) & ~3);
    const int b = blockIdx.x, lane = threadIdx.x;
    {  // stage the frame
        const uint8_t *src = D.scaled + (size_t)b * D.npx;
        if (!kImgSmem) {
            // nothing to stage
        } else if (((size_t)src & 15) == 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            for (int i = lane; i < D.npx / 16; i += 32) s_grow[i] = s4[i];
            for (int i = (D.npx / 16) * 16 + lane; i < D.npx; i += 32) s_img[i] = src[i];
        } else {
            for (int i = lane; i < D.npx; i += 32) s_img[i] = src[i];
        }
        for (int i = lane; i < used_words; i += 32) s_used[i] = 0;
    }
    __syncwarp();
    Grow G;
    G.sw = D.sw;
    G.sh = D.sh;
    G.kthr = D.kthr;
    G.density_th = D.density_th;
    G.img = kImgSmem ? s_img : D.scaled + (size_t)b * D.npx;
    G.reg_cap = kImgSmem ? kRegCap : D.reg_cap_small;
    G.used = s_used;
    G.mark = s_used;
    G.reg = s_reg;
    G.reg_ovf = D.reg_xy + (size_t)b * D.npx;
    G.tab = D.cstab;
    G.direct = (D.direct_trig & 2) != 0;
    G.claim = nullptr;
    G.aborted = false;
#ifdef PLP_LSD_PROF
    G.pc = s_pc;
#endif
    G.lane = lane;
    const uint32_t *order = D.order + (size_t)b * D.npx;
    float4 *segs = D.segs + (size_t)b * D.seg_cap;
    const int nseeds = D.nseeds[b];
    const int sw = D.sw;
    int nseg = 0;
    for (int s0 = 0; s0 < nseeds; s0 += 32) {
        const int s = s0 + lane;
        const uint32_t oxy = s < nseeds ? order[s] : 0;
        const int oidx = (int)(oxy >> 16) * sw + (int)(oxy & 0xffff);
        unsigned m = __ballot_sync(kFull, s < nseeds && !((s_used[oidx >> 5] >> (oidx & 31)) & 1u));
        while (m) {
            const int l = __ffs(m) - 1;
            const uint32_t seed_xy = __shfl_sync(kFull, oxy, l);
            const int sidx = __shfl_sync(kFull, oidx, l);
            int gx, gy;
            grad_at(G.img, sw, sidx, gx, gy);
            const float seed_deg = fast_atan2_deg((float)gx, (float)-gy);
            double reg_angle;
            PROF_T(t0);
            int n = region_grow(G, seed_xy, seed_deg, D.prec, reg_angle);
            PROF_ADD(0, t0);
#ifdef PLP_LSD_PROF
            cnt_regions++;
            cnt_px += n;
#endif
            if (n >= D.min_reg_size) {
                Rect R;
                PROF_T(t1);
                region2rect(G, n, reg_angle, D.prec, R);
                PROF_ADD(1, t1);
                PROF_T(t2);
                const bool okr = refine(G, n, seed_deg, reg_angle, D.prec, R);
                PROF_ADD(2, t2);
#ifdef PLP_LSD_PROF
                cnt_refine++;
#endif
                if (okr) {
                    if (nseg < D.seg_cap) {
                        if (lane == 0) {
                            // + 0.5 offset, then / scale (0.5)
                            segs[nseg] = make_float4((float)((R.x1 + 0.5) / 0.5), (float)((R.y1 + 0.5) / 0.5),
                                                     (float)((R.x2 + 0.5) / 0.5), (float)((R.y2 + 0.5) / 0.5));
                        }
                    } else if (lane == 0) {
                        atomicOr(&D.status[b], 1);
                    }
                    ++nseg;
                }
            }
            __syncwarp();
            m = __ballot_sync(kFull, s < nseeds && !((s_used[oidx >> 5] >> (oidx & 31)) & 1u)) & ~((2u << l) - 1u);
        }
    }
    if (lane == 0) D.nseg[b] = min(nseg, D.seg_cap);
#ifdef PLP_LSD_PROF
    if (lane == 0 && b == 0)
        printf("[lsd prof] total %lld grow %lld rect %lld refine %lld | seeds %d regions %lld px %lld big %lld segs %d | iters %lld rounds %lld load-cyc %lld resolve-cyc %lld ondemand %lld\n",
               clock64() - t_start, prof[0], prof[1], prof[2], nseeds, cnt_regions, cnt_px, cnt_refine, nseg, s_pc[0], s_pc[1], s_pc[2], s_pc[3], s_pc[4]);
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// K4': the same region growing for a SINGLE live frame (small batches): kMwWarps warps of one CTA work on consecutive
//      seeds of the frame speculatively and commit in seed order, so the result is the sequential one bit for bit.
//
// The sequential algorithm visits the seeds in order; what it does with a seed depends on the `used` map only through the
// pixels it ACCEPTS into a region (a neighbour that is not aligned is rejected whether it is used or not, a used one is
// never accepted).  One round: every warp takes the next not-yet-used seed (warp w the w-th), grows / refines it against
// the committed map plus a private mark bitmap, and records the bounding box of every pixel it accepted at any time.  The
// round's seeds e < w come earlier in the sequential order: if the box of w is disjoint from the boxes of all of them,
// nothing they mark can be a pixel w accepted, so w saw exactly the map the sequential run would have shown it.  The
// longest conflict-free prefix of the round commits (marks are OR-ed into the committed map, segments are emitted in
// seed order); the first conflicting seed and everything after it is redone in the next round, where it is first (and
// therefore commits): every round makes progress.  Marks of a committed region are final (refinement only ever clears a
// region's OWN pixels, before it commits), which is why skipping a seed that is used in the committed map is exact.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMwMaxWarps = 8;
constexpr int kMwRegCap = 1024;  // region entries per warp in shared memory (longer regions continue in global memory)

struct MwCtl {
    int bbox[kMwMaxWarps][4];
    int nfinal[kMwMaxWarps];
    int ok[kMwMaxWarps];
    float4 seg[kMwMaxWarps];
    unsigned long long stat[8];  // rounds, seeds run, seeds redone, cycles of warp 0: scan, own seed, wait, commit (tuning aid)
};

__global__ void __launch_bounds__(kMwMaxWarps * 32) lsd_grow_mw_kernel(LineDev D, uint32_t *reg_ovf_mw) {
    extern __shared__ uint4 s_grow[];
    const int W = blockDim.x >> 5;
    uint8_t *s_img = reinterpret_cast<uint8_t *>(s_grow);
    const int img_bytes = (D.npx + 15) & ~15;
    const int used_words = (D.npx + 31) >> 5, used_pad = (used_words + 3) & ~3;
    uint32_t *s_used = reinterpret_cast<uint32_t *>(s_img + img_bytes);  // committed marks
    uint32_t *s_priv = s_used + used_pad;                                // W private bitmaps
    uint32_t *s_reg = s_priv + (size_t)W * used_pad;                     // W region windows
    MwCtl &C = *reinterpret_cast<MwCtl *>(s_reg + (size_t)W * kMwRegCap);
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {  // stage the frame
        const uint8_t *src = D.scaled + (size_t)b * D.npx;
        if (((size_t)src & 15) == 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            for (int i = tid; i < D.npx / 16; i += blockDim.x) s_grow[i] = s4[i];
            for (int i = (D.npx / 16) * 16 + tid; i < D.npx; i += blockDim.x) s_img[i] = src[i];
        } else {
            for (int i = tid; i < D.npx; i += blockDim.x) s_img[i] = src[i];
        }
        for (int i = tid; i < used_pad * (W + 1); i += blockDim.x) s_used[i] = 0;
        if (tid < 8) C.stat[tid] = 0;
    }
    __syncthreads();
    GrowT<true> G;
    G.sw = D.sw;
    G.sh = D.sh;
    G.kthr = D.kthr;
    G.density_th = D.density_th;
    G.img = s_img;
    G.reg_cap = kMwRegCap;
    G.used = s_used;
    G.mark = s_priv + (size_t)warp * used_pad;
    G.reg = s_reg + (size_t)warp * kMwRegCap;
    G.reg_ovf = reg_ovf_mw + ((size_t)b * kMwMaxWarps + warp) * D.npx;
    G.tab = D.cstab;
    G.direct = (D.direct_trig & 1) != 0;
    G.claim = nullptr;
    G.aborted = false;
    G.lane = lane;
    const uint32_t *order = D.order + (size_t)b * D.npx;
    float4 *segs = D.segs + (size_t)b * D.seg_cap;
    const int nseeds = D.nseeds[b];
    const int sw = D.sw;
    int nseg = 0, cursor = 0;  // identical in every warp
    for (;;) {
        const long long tc0 = clock64();
        // ---- the next W seeds that are not used in the committed map (every warp scans for itself: same result)
        int my_pos = -1;       // lane i < W: position of the round's i-th seed in the order list
        uint32_t my_xy = 0;
        int found = 0, scan = cursor;
        while (found < W && scan < nseeds) {
            const int sidx_l = scan + lane;
            const uint32_t oxy = sidx_l < nseeds ? order[sidx_l] : 0;
            const int oidx = (int)(oxy >> 16) * sw + (int)(oxy & 0xffff);
            unsigned m = __ballot_sync(kFull, sidx_l < nseeds && !((s_used[oidx >> 5] >> (oidx & 31)) & 1u));
            while (m && found < W) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t xy = __shfl_sync(kFull, oxy, l);
                if (lane == found) {
                    my_pos = scan + l;
                    my_xy = xy;
                }
                ++found;
            }
            scan += 32;
        }
        if (found == 0) break;  // uniform over the CTA
        const long long tc1 = clock64();
        // ---- phase 1: warp w runs the w-th seed
        if (warp < found) {
            const uint32_t seed_xy = __shfl_sync(kFull, my_xy, warp);
            const int sidx = (int)(seed_xy >> 16) * sw + (int)(seed_xy & 0xffff);
            G.bx0 = G.by0 = 0x7fffffff;
            G.bx1 = G.by1 = -1;
            int gx, gy;
            grad_at(G.img, sw, sidx, gx, gy);
            const float seed_deg = fast_atan2_deg((float)gx, (float)-gy);
            double reg_angle;
            int n = region_grow(G, seed_xy, seed_deg, D.prec, reg_angle);
            bool okr = false;
            Rect R;
            if (n >= D.min_reg_size) {
                region2rect(G, n, reg_angle, D.prec, R);
                okr = refine(G, n, seed_deg, reg_angle, D.prec, R);
            }
            int x0 = G.bx0, y0 = G.by0, x1 = G.bx1, y1 = G.by1;
            for (int off = 16; off >= 1; off >>= 1) {
                x0 = min(x0, __shfl_xor_sync(kFull, x0, off));
                y0 = min(y0, __shfl_xor_sync(kFull, y0, off));
                x1 = max(x1, __shfl_xor_sync(kFull, x1, off));
                y1 = max(y1, __shfl_xor_sync(kFull, y1, off));
            }
            if (lane == 0) {
                C.bbox[warp][0] = x0;
                C.bbox[warp][1] = y0;
                C.bbox[warp][2] = x1;
                C.bbox[warp][3] = y1;
                C.nfinal[warp] = n;
                C.ok[warp] = okr ? 1 : 0;
                if (okr)  // + 0.5 offset, then / scale (0.5)
                    C.seg[warp] = make_float4((float)((R.x1 + 0.5) / 0.5), (float)((R.y1 + 0.5) / 0.5),
                                              (float)((R.x2 + 0.5) / 0.5), (float)((R.y2 + 0.5) / 0.5));
            }
        }
        const long long tc2 = clock64();
        __syncthreads();
        const long long tc3 = clock64();
        // ---- phase 2 (every warp, same result): first seed whose box meets the box of an earlier seed of the round
        int first_bad = found;
        {
            // lane = pair (v, e), e < v < found: at most 28 pairs
            int v = 1, e = lane;
            while (v < kMwMaxWarps && e >= v) {
                e -= v;
                ++v;
            }
            bool hit = false;
            if (v < found) {
                const int *bv = C.bbox[v], *be = C.bbox[e];
                hit = !(bv[2] < be[0] || be[2] < bv[0] || bv[3] < be[1] || be[3] < bv[1]);
            }
            for (int q = 1; q < found; ++q) {
                const unsigned mq = __ballot_sync(kFull, hit && v == q);
                if (mq && first_bad == found) first_bad = q;
            }
        }
        // ---- phase 3: commit the prefix, forget the rest
        if (warp < found) {
            const int n = C.nfinal[warp];
            const bool commit = warp < first_bad;
            for (int i = lane; i < n; i += 32) {
                const uint32_t xy = G.get(i);
                const int pidx = (int)(xy >> 16) * sw + (int)(xy & 0xffff);
                atomicAnd(&G.mark[pidx >> 5], ~(1u << (pidx & 31)));
                if (commit) atomicOr(&s_used[pidx >> 5], 1u << (pidx & 31));
            }
            if (commit && C.ok[warp] && lane == 0) {
                int slot = nseg;
                for (int e = 0; e < warp; ++e) slot += C.ok[e];
                if (slot < D.seg_cap) segs[slot] = C.seg[warp];
                else atomicOr(&D.status[b], 1);
            }
        }
        for (int e = 0; e < first_bad; ++e) nseg += C.ok[e];
        {
            const int pos_bad = __shfl_sync(kFull, my_pos, min(first_bad, found - 1));
            cursor = first_bad < found ? pos_bad : pos_bad + 1;
        }
        __syncthreads();
        if (tid == 0) {
            C.stat[0] += 1;
            C.stat[1] += (unsigned long long)found;
            C.stat[2] += (unsigned long long)(found - first_bad);
            C.stat[3] += (unsigned long long)(tc1 - tc0);
            C.stat[4] += (unsigned long long)(tc2 - tc1);
            C.stat[5] += (unsigned long long)(tc3 - tc2);
            C.stat[6] += (unsigned long long)(clock64() - tc3);
        }
    }
    if (tid == 0) {
        D.nseg[b] = min(nseg, D.seg_cap);
        if (D.mw_stat) {
            for (int q = 0; q < 7; ++q) D.mw_stat[8 * b + q] = C.stat[q];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K4'': region growing for one live frame, OUT OF ORDER with in-order commit (a reorder buffer, as in a CPU core).
//
// The round protocol above loses half of its cycles waiting for the slowest seed of a round: 72 % of the seeds grow fewer
// than 5 pixels, 10 % grow a few hundred, and a round costs its slowest member.  Here a warp that has finished a small
// region PARKS it (its <= kOooPark pixels and its bounding box go into the ticket's reorder-buffer entry, the private marks
// are cleared) and takes the next seed at once, so the warps that draw small seeds run ahead -- up to kOooWindow tickets --
// while the long regions of several lines are grown concurrently by the other warps (a region too large to park is HELD by
// its warp until it commits).  Tickets = seeds in gradient order.  A ticket runs against the committed `used` map plus its
// own private marks only; it records `start_head` = the commit pointer when it started.  Commit is strictly in ticket
// order, by whichever warp finds the head entry finished: the entry is valid if the bounding box of everything it accepted
// is disjoint from the final boxes of the tickets in [start_head, ticket) -- the tickets before start_head were committed,
// hence fully visible, when it started; the others can only matter through pixels it accepted (see K4').  An invalid entry,
// and an entry that was DEFERRED because its seed lay inside a region another warp was growing (most likely about to be
// absorbed), is simply executed AT THE HEAD, where the committed map is exactly the sequential state -- so every decision that
// is not provably the sequential one is redone sequentially, and the segments come out in ticket order, bit for bit.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kOooRing = 128;    // reorder-buffer entries (ticket % kOooRing)
constexpr int kOooWindow = 64;   // tickets in flight (<= kOooRing / 2: an entry is not reused while a later ticket may still test it)
constexpr int kOooPark = 24;     // pixels of a region that can be parked in its entry
constexpr int kOooRegCap = 512;  // region window per warp in shared memory (longer regions continue in global memory)

enum { kOooRunning = 1, kOooVoid = 2, kOooDeferred = 3, kOooParked = 4, kOooHeld = 5, kOooCommitted = 6 };

struct OooEntry {
    float4 seg;
    uint32_t xy;
    int state, start_head, owner, n, ok;
    short bx0, by0, bx1, by1;
    uint32_t px[kOooPark];
};

struct OooCtl {
    int lock_dispatch, lock_commit;
    int head, next_ticket;
    int scan_pos, chunk_base;
    unsigned chunk_mask;
    int nseg;
    uint32_t chunk_xy[32];
    int ctx_ticket[kMwMaxWarps + 1];  // ticket every context (warps + the head-execution context) works on, INT_MAX: none
    int pad_[3];
    unsigned long long stat[8];  // tickets, void, deferred, parked, held, executed at the head, conflicts, aborted
};

// (every wait of this kernel is bounded: after ~2 s of SM cycles a warp gives up, raises status bit 2 and leaves -- a protocol
// bug must not hang the device)
constexpr long long kOooTimeout = 4000000000ll;
__device__ __forceinline__ bool ooo_expired(long long t_start) {  // one lane decides (the lanes' clocks differ by a few cycles)
    int e = 0;
    if ((threadIdx.x & 31) == 0) e = clock64() - t_start > kOooTimeout ? 1 : 0;
    return __shfl_sync(kFull, e, 0) != 0;
}
__device__ __forceinline__ bool ooo_lock(int *l, int lane, long long t_start) {
    int ok = 1;
    if (lane == 0) {
        while (atomicCAS(l, 0, 1) != 0) {
            __nanosleep(40);
            if (clock64() - t_start > kOooTimeout) {
                ok = 0;
                break;
            }
        }
        __threadfence_block();
    }
    return __shfl_sync(kFull, ok, 0) != 0;
}
__device__ __forceinline__ bool ooo_trylock(int *l, int lane) {
    int got = 0;
    if (lane == 0) {
        got = atomicCAS(l, 0, 1) == 0 ? 1 : 0;
        if (got) __threadfence_block();
    }
    return __shfl_sync(kFull, got, 0) != 0;
}
__device__ __forceinline__ void ooo_unlock(int *l, int lane) {
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        atomicExch(l, 0);
    }
    __syncwarp();
}
// A word that another warp may change at any moment is read by ONE lane and broadcast: if every lane read it for itself the
// lanes of a warp could see different values and take different branches around warp-collective operations (this hung the
// first version of the kernel).
__device__ __forceinline__ int ooo_ld(const int *p) {
    int v = 0;
    if ((threadIdx.x & 31) == 0) v = *reinterpret_cast<const volatile int *>(p);
    return __shfl_sync(kFull, v, 0);
}

// one seed against `G.used | G.mark`: grow, rectangle, refinement.  Returns the final region size (marks left in G.mark, list in
// G.reg); ok / seg describe the segment; the bounding box of everything accepted is left in (x0, y0, x1, y1).
// (G by value: a private copy whose address never escapes, so that its fields -- the bounding box above all -- live in registers)
__device__ __forceinline__ int ooo_run_seed(const GrowT<true> G, const LineDev &D, uint32_t seed_xy, bool &okr, float4 &seg, int &x0,
                                            int &y0, int &x1, int &y1, bool &aborted) {
    const int sw = D.sw, lane = G.lane;
    const int sidx = (int)(seed_xy >> 16) * sw + (int)(seed_xy & 0xffff);
    G.bx0 = G.by0 = 0x7fffffff;
    G.bx1 = G.by1 = -1;
    int gx, gy;
    grad_at(G.img, sw, sidx, gx, gy);
    const float seed_deg = fast_atan2_deg((float)gx, (float)-gy);
    double reg_angle;
    G.aborted = false;
    int n = region_grow(G, seed_xy, seed_deg, D.prec, reg_angle);
    okr = false;
    Rect R;
    if (!G.aborted && n >= D.min_reg_size) {
        region2rect(G, n, reg_angle, D.prec, R);
        okr = refine(G, n, seed_deg, reg_angle, D.prec, R);
    }
    aborted = G.aborted;
    if (aborted) okr = false;
    x0 = G.bx0, y0 = G.by0, x1 = G.bx1, y1 = G.by1;
    for (int off = 16; off >= 1; off >>= 1) {
        x0 = min(x0, __shfl_xor_sync(kFull, x0, off));
        y0 = min(y0, __shfl_xor_sync(kFull, y0, off));
        x1 = max(x1, __shfl_xor_sync(kFull, x1, off));
        y1 = max(y1, __shfl_xor_sync(kFull, y1, off));
    }
    seg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (okr)  // + 0.5 offset, then / scale (0.5)
        seg = make_float4((float)((R.x1 + 0.5) / 0.5), (float)((R.y1 + 0.5) / 0.5), (float)((R.x2 + 0.5) / 0.5),
                          (float)((R.y2 + 0.5) / 0.5));
    (void)lane;
    return n;
}

struct OooShared {
    OooCtl *C;
    OooEntry *ring;
    uint32_t *used;  // committed marks
    int used_pad, sw;
};

// region list of G (n entries) -> committed map; the private marks are cleared
__device__ __forceinline__ void ooo_commit_list(const GrowT<true> &G, uint32_t *used, int n, int sw, bool commit) {
    for (int i = G.lane; i < n; i += 32) {
        const uint32_t xy = G.get(i);
        const int pidx = (int)(xy >> 16) * sw + (int)(xy & 0xffff);
        atomicAnd(&G.mark[pidx >> 5], ~(1u << (pidx & 31)));
        G.unclaim(pidx);
        if (commit) atomicOr(&used[pidx >> 5], 1u << (pidx & 31));
    }
    __syncwarp();
}

// Commit finished tickets from the head, in order (warp-collective; returns at once if another warp is draining).
// `Gown`: the caller's context (its HELD region, if any, is committed from it); `Gsp`: the spare context that whoever holds the
// commit lock uses to execute a ticket at the head.
__device__ void ooo_drain(const OooShared &S, const LineDev &D, const GrowT<true> &Gown, const GrowT<true> &Gsp, int warp, float4 *segs,
                          int b) {
    OooCtl &C = *S.C;
    const int lane = Gown.lane;
    if (!ooo_trylock(&C.lock_commit, lane)) return;
    for (;;) {
        const int h = ooo_ld(&C.head);
        if (h == ooo_ld(&C.next_ticket)) break;
        OooEntry &E = S.ring[h & (kOooRing - 1)];
        const int st = ooo_ld(&E.state);
        if (st == kOooRunning) break;
        const int owner = ooo_ld(&E.owner);
        if (st == kOooHeld && owner != warp) break;  // its owner commits it (it is spinning on this lock)
        bool exec = st == kOooDeferred;
        if (st == kOooParked || st == kOooHeld) {
            // valid  <=>  box disjoint from the final boxes of the tickets that were not yet committed when it started
            bool hit = false;
            const int bx0 = E.bx0, by0 = E.by0, bx1 = E.bx1, by1 = E.by1;
            for (int e = E.start_head + lane; e < h; e += 32) {
                const OooEntry &F = S.ring[e & (kOooRing - 1)];
                hit = hit || !(F.bx1 < bx0 || bx1 < F.bx0 || F.by1 < by0 || by1 < F.by0);  // (an empty box has bx1 = -1 < bx0)
            }
            if (__any_sync(kFull, hit)) {
                exec = true;
                if (lane == 0) C.stat[6]++;
            }
        }
        int nseg = ooo_ld(&C.nseg);
        if (st == kOooVoid) {
            // nothing
        } else if (exec) {
            if (st == kOooHeld) ooo_commit_list(Gown, S.used, E.n, S.sw, false);  // drop the caller's speculative region
            const uint32_t seed_xy = E.xy;
            const int sidx = (int)(seed_xy >> 16) * S.sw + (int)(seed_xy & 0xffff);
            int n = 0, x0 = 0x7fffffff, y0 = 0x7fffffff, x1 = -1, y1 = -1;
            bool okr = false;
            float4 seg = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!((ooo_ld(reinterpret_cast<const int *>(&S.used[sidx >> 5])) >> (sidx & 31)) & 1)) {  // the committed map IS the sequential state here
                GrowT<true> Gh = Gsp;
                Gh.my_ticket = h;  // the lowest ticket in flight: nothing it meets can belong to an earlier one, it never stops
                if (lane == 0) *reinterpret_cast<volatile int *>(&C.ctx_ticket[Gsp.self]) = h;
                __syncwarp();
                bool ab;
                n = ooo_run_seed(Gh, D, seed_xy, okr, seg, x0, y0, x1, y1, ab);
                ooo_commit_list(Gsp, S.used, n, S.sw, true);
                if (lane == 0) *reinterpret_cast<volatile int *>(&C.ctx_ticket[Gsp.self]) = 0x7fffffff;
            }
            if (lane == 0) {
                E.bx0 = (short)min(x0, 32767);
                E.by0 = (short)min(y0, 32767);
                E.bx1 = (short)x1;
                E.by1 = (short)y1;
                C.stat[5]++;
                if (okr) {
                    if (nseg < D.seg_cap) segs[nseg] = seg;
                    else atomicOr(&D.status[b], 1);
                    C.nseg = nseg + 1;
                }
            }
        } else {  // a valid speculative result
            if (st == kOooParked) {
                if (lane < E.n) {
                    const uint32_t xy = E.px[lane];
                    const int pidx = (int)(xy >> 16) * S.sw + (int)(xy & 0xffff);
                    atomicOr(&S.used[pidx >> 5], 1u << (pidx & 31));
                }
            } else {
                ooo_commit_list(Gown, S.used, E.n, S.sw, true);
            }
            if (lane == 0 && E.ok) {
                if (nseg < D.seg_cap) segs[nseg] = E.seg;
                else atomicOr(&D.status[b], 1);
                C.nseg = nseg + 1;
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (st == kOooHeld) *reinterpret_cast<volatile int *>(&E.state) = kOooCommitted;  // releases the owner (the caller)
            *reinterpret_cast<volatile int *>(&C.head) = h + 1;
        }
        __syncwarp();
    }
    ooo_unlock(&C.lock_commit, lane);
}

__global__ void __launch_bounds__(kMwMaxWarps * 32) lsd_grow_ooo_kernel(LineDev D, uint32_t *reg_ovf_mw) {
    extern __shared__ uint4 s_grow[];
    const int W = blockDim.x >> 5;
    uint8_t *s_img = reinterpret_cast<uint8_t *>(s_grow);
    const int img_bytes = (D.npx + 15) & ~15;
    const int used_words = (D.npx + 31) >> 5, used_pad = (used_words + 3) & ~3;
    uint32_t *s_used = reinterpret_cast<uint32_t *>(s_img + img_bytes);  // committed marks
    uint32_t *s_priv = s_used + used_pad;                                // W + 1 private bitmaps (the last one: head execution)
    uint32_t *s_claim = s_priv + (size_t)(W + 1) * used_pad;             // union of the private bitmaps (collision detector)
    uint32_t *s_reg = s_claim + used_pad;                                // W + 1 region windows
    OooEntry *ring = reinterpret_cast<OooEntry *>(s_reg + (size_t)(W + 1) * kOooRegCap);
    OooCtl &C = *reinterpret_cast<OooCtl *>(ring + kOooRing);
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {  // stage the frame
        const uint8_t *src = D.scaled + (size_t)b * D.npx;
        if (((size_t)src & 15) == 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            for (int i = tid; i < D.npx / 16; i += blockDim.x) s_grow[i] = s4[i];
            for (int i = (D.npx / 16) * 16 + tid; i < D.npx; i += blockDim.x) s_img[i] = src[i];
        } else {
            for (int i = tid; i < D.npx; i += blockDim.x) s_img[i] = src[i];
        }
        for (int i = tid; i < used_pad * (W + 3); i += blockDim.x) s_used[i] = 0;
        uint32_t *z = reinterpret_cast<uint32_t *>(ring);
        for (int i = tid; i < (int)((sizeof(OooEntry) * kOooRing + sizeof(OooCtl)) / 4); i += blockDim.x) z[i] = 0;
    }
    __syncthreads();
    GrowT<true> G, Gsp;
    G.sw = D.sw;
    G.sh = D.sh;
    G.kthr = D.kthr;
    G.density_th = D.density_th;
    G.img = s_img;
    G.reg_cap = kOooRegCap;
    G.used = s_used;
    G.tab = D.cstab;
    G.direct = false;
    G.lane = lane;
    G.claim = s_claim;
    G.priv_base = s_priv;
    G.ctx_ticket = C.ctx_ticket;
    G.nctx = W + 1;
    G.ctx_words = used_pad;
    G.my_ticket = 0x7fffffff;
    G.aborted = false;
    G.self = warp;
    Gsp = G;
    Gsp.self = W;
    G.mark = s_priv + (size_t)warp * used_pad;
    G.reg = s_reg + (size_t)warp * kOooRegCap;
    G.reg_ovf = reg_ovf_mw + ((size_t)b * (kMwMaxWarps + 1) + warp) * D.npx;
    Gsp.mark = s_priv + (size_t)W * used_pad;
    Gsp.reg = s_reg + (size_t)W * kOooRegCap;
    Gsp.reg_ovf = reg_ovf_mw + ((size_t)b * (kMwMaxWarps + 1) + kMwMaxWarps) * D.npx;
    OooShared S{&C, ring, s_used, used_pad, D.sw};
    const uint32_t *order = D.order + (size_t)b * D.npx;
    float4 *segs = D.segs + (size_t)b * D.seg_cap;
    const int nseeds = D.nseeds[b];
    const int sw = D.sw;
    if (tid <= W) C.ctx_ticket[tid] = 0x7fffffff;
    __syncthreads();
    const long long t_start = clock64();
    bool timed_out = false;
    for (;;) {
        if (ooo_expired(t_start)) {
            timed_out = true;
            break;
        }
        // ---- take the next ticket: the next seed (in order) that is not used in the committed map
        int t = -1;  // -1: no seed left, -2: the window is full
        if (!ooo_lock(&C.lock_dispatch, lane, t_start)) {
            timed_out = true;
            break;
        }
        {
            const int nt = ooo_ld(&C.next_ticket);
            if (nt - ooo_ld(&C.head) >= kOooWindow) {
                t = -2;
            } else {
                unsigned m = (unsigned)ooo_ld(reinterpret_cast<const int *>(&C.chunk_mask));
                int base = ooo_ld(&C.chunk_base), sp = ooo_ld(&C.scan_pos);
                uint32_t cxy = *reinterpret_cast<volatile uint32_t *>(&C.chunk_xy[lane]);
                while (m == 0u && sp < nseeds) {
                    const int p = sp + lane;
                    cxy = p < nseeds ? order[p] : 0u;
                    const int oidx = (int)(cxy >> 16) * sw + (int)(cxy & 0xffff);
                    m = __ballot_sync(kFull, p < nseeds && !((s_used[oidx >> 5] >> (oidx & 31)) & 1u));
                    base = sp;
                    sp += 32;
                }
                if (m != 0u) {
                    const int l = __ffs(m) - 1;
                    const uint32_t xy = __shfl_sync(kFull, cxy, l);
                    const int head_now = ooo_ld(&C.head);
                    t = nt;
                    if (lane == 0) {
                        OooEntry &E = ring[t & (kOooRing - 1)];
                        E.xy = xy;
                        E.start_head = head_now;
                        E.owner = warp;
                        E.n = 0;
                        E.ok = 0;
                        E.bx0 = E.by0 = 32767;
                        E.bx1 = E.by1 = -1;
                        *reinterpret_cast<volatile int *>(&E.state) = kOooRunning;
                        C.stat[0]++;
                    }
                    m &= m - 1;
                }
                C.chunk_xy[lane] = cxy;
                if (lane == 0) {
                    C.chunk_mask = m;
                    C.chunk_base = base;
                    C.scan_pos = sp;
                    if (t >= 0) {
                        __threadfence_block();
                        *reinterpret_cast<volatile int *>(&C.next_ticket) = t + 1;
                    }
                }
            }
        }
        ooo_unlock(&C.lock_dispatch, lane);
        if (t < 0) {
            if (t == -1 && ooo_ld(&C.head) == ooo_ld(&C.next_ticket)) {
                // every ticket is committed; a seed can only have been left behind if another warp is between its scan and
                // its ticket, which the dispatch lock excludes
                break;
            }
            ooo_drain(S, D, G, Gsp, warp, segs, b);
            __nanosleep(100);
            continue;
        }
        OooEntry &E = ring[t & (kOooRing - 1)];
        const uint32_t seed_xy = E.xy;
        const int sidx = (int)(seed_xy >> 16) * sw + (int)(seed_xy & 0xffff);
        int state;
        // (the committed map changes under our feet: one lane reads the word, see ooo_ld)
        if ((ooo_ld(reinterpret_cast<const int *>(&s_used[sidx >> 5])) >> (sidx & 31)) & 1) {
            state = kOooVoid;  // committed since the scan: by a ticket before this one, so the sequential run skips it too
        } else {
            // inside a region another warp is growing right now: most likely absorbed -- decided at the head instead
            bool other = false;
            if (lane < W && lane != warp) other = (s_priv[(size_t)lane * used_pad + (sidx >> 5)] >> (sidx & 31)) & 1u;
            if (__any_sync(kFull, other)) {
                state = kOooDeferred;
            } else {
                bool okr, aborted;
                float4 seg;
                int x0, y0, x1, y1;
                G.my_ticket = t;
                if (lane == 0) *reinterpret_cast<volatile int *>(&C.ctx_ticket[warp]) = t;
                __syncwarp();
                const int n = ooo_run_seed(G, D, seed_xy, okr, seg, x0, y0, x1, y1, aborted);
                if (aborted) {
                    // ran into a pixel of an earlier region in flight: most likely this seed is about to be absorbed -- forget the
                    // partial region and let the head decide
                    ooo_commit_list(G, s_used, n, sw, false);
                    if (lane == 0) {
                        *reinterpret_cast<volatile int *>(&C.ctx_ticket[warp]) = 0x7fffffff;
                        C.stat[7]++;
                    }
                    state = kOooDeferred;
                } else {
                if (lane == 0) {
                    E.n = n;
                    E.ok = okr ? 1 : 0;
                    E.seg = seg;
                    E.bx0 = (short)x0;
                    E.by0 = (short)y0;
                    E.bx1 = (short)x1;
                    E.by1 = (short)y1;
                }
                if (n <= kOooPark) {
                    if (lane < n) E.px[lane] = G.get(lane);
                    __syncwarp();
                    ooo_commit_list(G, s_used, n, sw, false);  // clears the private marks only
                    if (lane == 0) *reinterpret_cast<volatile int *>(&C.ctx_ticket[warp]) = 0x7fffffff;
                    state = kOooParked;
                } else {
                    state = kOooHeld;  // (the context keeps its ticket until the region is committed)
                }
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            C.stat[state - 1]++;  // (counted without the lock: a tuning aid, may lose increments)
            __threadfence_block();
            *reinterpret_cast<volatile int *>(&E.state) = state;
        }
        __syncwarp();
        if (state == kOooHeld) {
            while (ooo_ld(&E.state) == kOooHeld) {
                ooo_drain(S, D, G, Gsp, warp, segs, b);
                if (ooo_ld(&E.state) == kOooHeld) __nanosleep(100);
                if (ooo_expired(t_start)) {
                    timed_out = true;
                    break;
                }
            }
            if (timed_out) break;
            if (lane == 0) *reinterpret_cast<volatile int *>(&C.ctx_ticket[warp]) = 0x7fffffff;
        } else {
            ooo_drain(S, D, G, Gsp, warp, segs, b);
        }
    }
    if (timed_out && lane == 0) atomicOr(&D.status[b], 2);
    __syncthreads();
    if (tid == 0) {
        D.nseg[b] = min(C.nseg, D.seg_cap);
        if (D.mw_stat)
            for (int q = 0; q < 8; ++q) D.mw_stat[8 * b + q] = C.stat[q];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K5: segments -> KeyLines (LSDDetector_custom.cpp:266-300) + 2-D line functions (line_extractor.cc:147-159)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) keyline_kernel(LineDev D, plp_keyline *kl_out, double *fn_out, int32_t *n_out) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const float4 *segs = D.segs + (size_t)b * D.seg_cap;
    plp_keyline *kls = kl_out + (size_t)b * D.kl_cap;
    double *fns = fn_out + (size_t)b * D.kl_cap * 3;
    const int nseg = D.nseg[b];
    int nk = 0, nc = 0;
    for (int s0 = 0; s0 < nseg; s0 += 32) {
        const int s = s0 + lane;
        float e0 = 0, e1 = 0, e2 = 0, e3 = 0;
        double length = 0;
        bool pass = false;
        if (s < nseg) {
            const float4 v = segs[s];
            e0 = v.x; e1 = v.y; e2 = v.z; e3 = v.w;
            // checkLineExtremes
            if (e0 < 0) e0 = 0;
            if (e0 >= D.w) e0 = (float)D.w - 1.0f;
            if (e2 < 0) e2 = 0;
            if (e2 >= D.w) e2 = (float)D.w - 1.0f;
            if (e1 < 0) e1 = 0;
            if (e1 >= D.h) e1 = (float)D.h - 1.0f;
            if (e3 < 0) e3 = 0;
            if (e3 >= D.h) e3 = (float)D.h - 1.0f;
            const double ddx = (double)(e0 - e2), ddy = (double)(e1 - e3);
            length = (double)(float)sqrt(ddx * ddx + ddy * ddy);
            // LSDDetector_custom.cpp:270 length > min_length; line_extractor.cc:136 lineLength >= 60
            pass = (length > D.min_length) && ((float)length >= 60.f);
        }
        // class_id counts every segment with length > min_length (also those the >= 60 filter would drop); with
        // min_length = 0.125 * min(w, h) >= 60 both filters coincide for images of at least 480 rows
        const bool counted = (s < nseg) && (length > D.min_length);
        const unsigned cm = __ballot_sync(kFull, counted);
        const unsigned pm = __ballot_sync(kFull, pass);
        if (pass) {
            const int pos = nk + __popc(pm & ((1u << lane) - 1));
            if (pos < D.kl_cap) {
                plp_keyline k;
                k.start_x = e0; k.start_y = e1; k.end_x = e2; k.end_y = e3;
                k.s_oct_x = e0; k.s_oct_y = e1; k.e_oct_x = e2; k.e_oct_y = e3;
                k.line_length = (float)length;
                const int x0 = __float2int_rn(e0), y0 = __float2int_rn(e1), x1 = __float2int_rn(e2), y1 = __float2int_rn(e3);
                k.num_pixels = max(abs(x1 - x0), abs(y1 - y0)) + 1;
                const float ay = e3 - e1, ax = e2 - e0;
                k.angle = (float)det_atan2((double)ay, (double)ax);
                k.class_id = nc + __popc(cm & ((1u << lane) - 1));
                k.octave = 0;
                k.size = (e2 - e0) * (e3 - e1);
                k.response = k.line_length / (float)max(D.w, D.h);
                k.pt_x = (e2 + e0) / 2;
                k.pt_y = (e3 + e1) / 2;
                kls[pos] = k;
                const double sx = e0, sy = e1, ex = e2, ey = e3;
                const double l0 = sy - ey, l1 = ex - sx, l2 = sx * ey - sy * ex;
                const double nrm = sqrt(l0 * l0 + l1 * l1);
                fns[3 * pos] = l0 / nrm;
                fns[3 * pos + 1] = l1 / nrm;
                fns[3 * pos + 2] = l2 / nrm;
            } else {
                atomicOr(&D.status[b], 2);
            }
        }
        nk += __popc(pm);
        nc += __popc(cm);
    }
    if (lane == 0) n_out[b] = min(nk, D.kl_cap);
}

// ------------------------------------------------------------------------------------------------------------------
// K6: GaussianBlur(5x5, sigma 1) (Q8 taps 14 62 104 62 14) + Sobel 3x3 -> int16 (dx, dy); reflect-101 borders
//     (binary_descriptor_custom.cpp:347-395).  One CTA = 32 x 8 pixels.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGrTw = 32, kGrTh = 8;
__global__ void __launch_bounds__(256) lbd_gradient_kernel(LineDev D) {
    // The reflect-101 extension of the image is symmetric about every border, and so is its blur with a symmetric
    // kernel: blurred(reflect(p)) == blur of the extended image at p.  The tile is therefore staged by plain reflected
    // coordinates (blur halo 2 + Sobel halo 1) and filtered separably.
    __shared__ uint8_t s_src[kGrTh + 6][kGrTw + 8];
    __shared__ uint16_t s_h[kGrTh + 6][kGrTw + 2];
    __shared__ uint8_t s_b[kGrTh + 2][kGrTw + 2];
    const int b = blockIdx.y;
    const int tiles_x = (D.w + kGrTw - 1) / kGrTw;
    const int tx = (blockIdx.x % tiles_x) * kGrTw, ty = (blockIdx.x / tiles_x) * kGrTh;
    const uint8_t *img = D.img + (size_t)b * D.img_frame_stride;
    for (int i = threadIdx.x; i < (kGrTh + 6) * (kGrTw + 6); i += blockDim.x) {
        const int r = i / (kGrTw + 6), c = i - r * (kGrTw + 6);
        s_src[r][c] = img[(size_t)reflect101(ty + r - 3, D.h) * D.img_step + reflect101(tx + c - 3, D.w)];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (kGrTh + 6) * (kGrTw + 2); i += blockDim.x) {
        const int r = i / (kGrTw + 2), c = i - r * (kGrTw + 2);
        const uint8_t *q = &s_src[r][c];
        s_h[r][c] = (uint16_t)(14 * (q[0] + q[4]) + 62 * (q[1] + q[3]) + 104 * q[2]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (kGrTh + 2) * (kGrTw + 2); i += blockDim.x) {
        const int r = i / (kGrTw + 2), c = i - r * (kGrTw + 2);
        const uint32_t v = 14u * (s_h[r][c] + s_h[r + 4][c]) + 62u * (s_h[r + 1][c] + s_h[r + 3][c]) + 104u * s_h[r + 2][c];
        s_b[r][c] = (uint8_t)((v + 32768u) >> 16);
    }
    __syncthreads();
    short2 *out = D.grad + (size_t)b * D.w * D.h;
    for (int i = threadIdx.x; i < kGrTh * kGrTw; i += blockDim.x) {
        const int r = i / kGrTw, c = i - r * kGrTw;
        const int x = tx + c, y = ty + r;
        if (x >= D.w || y >= D.h) continue;
        const int a00 = s_b[r][c], a01 = s_b[r][c + 1], a02 = s_b[r][c + 2];
        const int a10 = s_b[r + 1][c], a12 = s_b[r + 1][c + 2];
        const int a20 = s_b[r + 2][c], a21 = s_b[r + 2][c + 1], a22 = s_b[r + 2][c + 2];
        const int gx = (a02 - a00) + 2 * (a12 - a10) + (a22 - a20);
        const int gy = (a20 - a00) + 2 * (a21 - a01) + (a22 - a02);
        out[(size_t)y * D.w + x] = make_short2((short)gx, (short)gy);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K7: LBD (binary_descriptor_custom.cpp:1018-1364 + 398-408, 642-646): one CTA of 64 threads per line; thread = row
//     of the 63-row line support region (the row sums are sequential float accumulations along the line, kept in the
//     reference's order); 9 threads accumulate the bands in row order; thread 0 normalises and packs the 32 bytes.
// ------------------------------------------------------------------------------------------------------------------
__constant__ int c_comb[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
                                  {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
                                  {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

__global__ void __launch_bounds__(64) lbd_kernel(LineDev D, const plp_keyline *kl_in, const int32_t *n_in, uint8_t *lbd_out) {
    __shared__ float s_row[kLspHeight][4];   // pl, nl, po, no of each row (already multiplied by the global weight)
    __shared__ float s_band[8][kBands];
    __shared__ float s_des[kBands * 8];
    const int b = blockIdx.y, t = threadIdx.x;
    const int n = n_in[b];
    const short2 *grad = D.grad + (size_t)b * D.w * D.h;
    for (int line = blockIdx.x; line < n; line += gridDim.x) {
        const plp_keyline kl = kl_in[(size_t)b * D.kl_cap + line];
        const short image_w = (short)(D.w - 1), image_h = (short)(D.h - 1);
        const short length_lsp = (short)kl.num_pixels;
        const short half_h = (kLspHeight - 1) / 2;
        const short half_w = (length_lsp - 1) / 2;
        const float mid_x = (float)(0.5 * (kl.s_oct_x + kl.e_oct_x));
        const float mid_y = (float)(0.5 * (kl.s_oct_y + kl.e_oct_y));
        const float dl0 = (float)det_cos((double)kl.angle), dl1 = (float)det_sin((double)kl.angle);
        const float do0 = -dl1, do1 = dl0;
        if (t < kLspHeight) {
            float scx0 = -dl0 * half_w + dl1 * half_h + mid_x;
            float scy0 = -dl1 * half_w - dl0 * half_h + mid_y;
            for (int hh = 0; hh < t; ++hh) {  // the reference walks the rows with running float sums
                scx0 -= dl1;
                scy0 += dl0;
            }
            float scx = scx0, scy = scy0;
            float pl = 0, nl = 0, po = 0, no = 0;
            for (short wid = 0; wid < length_lsp; ++wid) {
                short tc = (short)roundf(scx);
                const short xc = (tc < 0) ? 0 : (tc > image_w) ? image_w : tc;
                tc = (short)roundf(scy);
                const short yc = (tc < 0) ? 0 : (tc > image_h) ? image_h : tc;
                const short2 gd = grad[(size_t)yc * D.w + xc];
                const float gdl = gd.x * dl0 + gd.y * dl1;
                const float gdo = gd.x * do0 + gd.y * do1;
                if (gdl > 0) pl += gdl; else nl -= gdl;
                if (gdo > 0) po += gdo; else no -= gdo;
                scx += dl0;
                scy += dl1;
            }
            const float coef = D.gauss_g[t];
            s_row[t][0] = coef * pl;
            s_row[t][1] = coef * nl;
            s_row[t][2] = coef * po;
            s_row[t][3] = coef * no;
        }
        __syncthreads();
        if (t < kBands) {
            float bs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const int h0 = max(0, (t - 1) * kBandWidth), h1 = min(kLspHeight, (t + 2) * kBandWidth);
            for (int hid = h0; hid < h1; ++hid) {
                const int hb = hid / kBandWidth;
                // row of band hb contributes to band t with: own band -> gl[r + 7]; band above (t == hb - 1) -> gl[r + 14];
                // band below (t == hb + 1) -> gl[r]
                const int rr = hid % kBandWidth;
                const float c = (t == hb) ? D.gauss_l[rr + kBandWidth] : (t == hb - 1) ? D.gauss_l[rr + 2 * kBandWidth] : D.gauss_l[rr];
                const float pl = s_row[hid][0], nl = s_row[hid][1], po = s_row[hid][2], no = s_row[hid][3];
                const float pl2 = pl * pl, nl2 = nl * nl, po2 = po * po, no2 = no * no;
                bs[0] += c * pl;
                bs[1] += c * nl;
                bs[2] += c * c * pl2;
                bs[3] += c * c * nl2;
                bs[4] += c * po;
                bs[5] += c * no;
                bs[6] += c * c * po2;
                bs[7] += c * c * no2;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) s_band[q][t] = bs[q];
        }
        __syncthreads();
        if (t == 0) {
            const float inv_n2 = (float)(1.0 / (kBandWidth * 2.0)), inv_n3 = (float)(1.0 / (kBandWidth * 3.0));
            for (int bb = 0; bb < kBands; ++bb) {
                const float inv_n = (bb == 0 || bb == kBands - 1) ? inv_n2 : inv_n3;
                float *d = s_des + bb * 8;
                float tt = s_band[0][bb] * inv_n;
                d[0] = tt;
                d[4] = sqrtf(s_band[2][bb] * inv_n - tt * tt);
                tt = s_band[1][bb] * inv_n;
                d[1] = tt;
                d[5] = sqrtf(s_band[3][bb] * inv_n - tt * tt);
                tt = s_band[4][bb] * inv_n;
                d[2] = tt;
                d[6] = sqrtf(s_band[6][bb] * inv_n - tt * tt);
                tt = s_band[5][bb] * inv_n;
                d[3] = tt;
                d[7] = sqrtf(s_band[7][bb] * inv_n - tt * tt);
            }
            float tm = 0, ts = 0;
            for (int bb = 0; bb < kBands; ++bb) {
                const float *d = s_des + bb * 8;
                tm += d[0] * d[0];
                tm += d[1] * d[1];
                tm += d[2] * d[2];
                tm += d[3] * d[3];
                ts += d[4] * d[4];
                ts += d[5] * d[5];
                ts += d[6] * d[6];
                ts += d[7] * d[7];
            }
            tm = 1 / sqrtf(tm);
            ts = 1 / sqrtf(ts);
            for (int bb = 0; bb < kBands; ++bb) {
                float *d = s_des + bb * 8;
                for (int q = 0; q < 4; ++q) d[q] = d[q] * tm;
                for (int q = 4; q < 8; ++q) d[q] = d[q] * ts;
            }
            for (int q = 0; q < kBands * 8; ++q)
                if ((double)s_des[q] > 0.4) s_des[q] = (float)0.4;
            float tt = 0;
            for (int q = 0; q < kBands * 8; ++q) tt += s_des[q] * s_des[q];
            tt = 1 / sqrtf(tt);
            for (int q = 0; q < kBands * 8; ++q) s_des[q] = s_des[q] * tt;
        }
        __syncthreads();
        if (t < 32) {
            const float *f1 = s_des + 8 * c_comb[t][0], *f2 = s_des + 8 * c_comb[t][1];
            unsigned r = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (f1[q] > f2[q]) r |= (1u << q);
            lbd_out[((size_t)b * D.kl_cap + line) * 32 + t] = (uint8_t)r;
        }
        if (D.lbd_float)
            for (int q = t; q < kBands * 8; q += 64) D.lbd_float[((size_t)b * D.kl_cap + line) * 72 + q] = s_des[q];
        __syncthreads();
    }
}

}  // namespace

// ====================================================================================================================
struct plp_line {
    plp_ctx *ctx = nullptr;
    int rows = 0, cols = 0, max_batch = 0, last_batch = 0;
    LineDev dev{};
    uint8_t *d_img = nullptr;
    plp_keyline *d_kl = nullptr;
    uint8_t *d_lbd = nullptr;
    double *d_fn = nullptr;
    int32_t *d_n = nullptr;
    size_t sort_smem = 0, grow_smem = 0, grow_smem_noimg = 0;
    bool img_smem_ok = true;
    int resident_smem_frames = 0;
    bool force_global_image = false;
    int grow_variant = 0;  // 0 automatic, 1 one warp per frame, 2 multi-warp rounds (lsd_grow_mw_kernel), 3 out of order (lsd_grow_ooo_kernel)
    int ooo_warps = 0;
    bool ooo_auto = false;
    size_t ooo_smem = 0;
    int mw_warps = 0, mw_max_batch = 0;
    size_t mw_smem = 0;
    uint32_t *d_reg_mw = nullptr;
    float4 *d_cstab = nullptr;
    std::vector<void *> owned;
};

template <class T>
static plp_status dev_alloc(plp_line *h, T **p, size_t count) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
        return PLP_ERR_CUDA;
    }
    h->owned.push_back(q);
    *p = (T *)q;
    return PLP_OK;
}

static plp_status line_run(plp_line *h, const uint8_t *d_imgs, int batch, size_t step, plp_keyline *d_kl, uint8_t *d_lbd,
                           double *d_fn, int32_t *d_n, int32_t *d_status) {
    plp_ctx *ctx = h->ctx;
    LineDev D = h->dev;
    D.img = d_imgs;
    D.img_step = step;
    D.img_frame_stride = step * (size_t)h->rows;
    if (d_status) D.status = d_status;
    h->last_batch = batch;
    PLP_CUDA_TRY(cudaMemsetAsync(D.status, 0, (size_t)batch * sizeof(int), ctx->stream));
    {
        dim3 grid(div_up(D.w, kScTw) * div_up(D.h, kScTh), batch);
        PLP_LAUNCH(ctx, lsd_scale_kernel, grid, 256, 0, D);
    }
    PLP_LAUNCH(ctx, lsd_sort_kernel, batch, kSortWarps * 32, h->sort_smem, D);
    // small batches (at most half of what stays resident, so that a second handle -- the right image of a stereo pair --
    // fits beside it): image in shared memory (latency); larger batches: image through L2, 3x the frames per SM
    // a wave of frames or less: the frame-level parallelism cannot fill the GPU, several warps per frame (speculative, in-order
    // commit) cut the latency of a live frame instead
    const bool mw = h->mw_warps >= 2 && batch <= h->mw_max_batch && h->grow_variant != 1 &&
                    (h->grow_variant >= 2 || 2 * batch <= ctx->sm_count);  // half a wave: a second handle (stereo) fits beside it
    // (automatic mode takes the out-of-order kernel only when PLP_LSD_OOO=1: it is the newest code of the tree)
    const bool ooo = mw && h->ooo_warps >= 2 && (h->grow_variant == 3 || (h->grow_variant == 0 && h->ooo_auto));
    if (ooo) {
        PLP_LAUNCH(ctx, lsd_grow_ooo_kernel, batch, h->ooo_warps * 32, h->ooo_smem, D, h->d_reg_mw);
    } else if (mw) {
        PLP_LAUNCH(ctx, lsd_grow_mw_kernel, batch, h->mw_warps * 32, h->mw_smem, D, h->d_reg_mw);
    } else if (h->img_smem_ok && 2 * batch <= h->resident_smem_frames && !h->force_global_image) {
        PLP_LAUNCH(ctx, lsd_grow_kernel<true>, batch, 32, h->grow_smem, D);
    } else {
        PLP_LAUNCH(ctx, lsd_grow_kernel<false>, batch, 32, h->grow_smem_noimg, D);
    }
    PLP_LAUNCH(ctx, keyline_kernel, batch, 32, 0, D, d_kl, d_fn, d_n);
    {
        dim3 grid(div_up(D.w, kGrTw) * div_up(D.h, kGrTh), batch);
        PLP_LAUNCH(ctx, lbd_gradient_kernel, grid, 256, 0, D);
    }
    {
        dim3 grid(256, batch);
        PLP_LAUNCH(ctx, lbd_kernel, grid, 64, 0, D, d_kl, d_n, d_lbd);
    }
    PLP_CHECK_LAUNCH();
    return PLP_OK;
}

extern "C" {

void plp_line_destroy(plp_line *h) {
    if (!h) return;
    cudaSetDevice(h->ctx->device);
    cudaStreamSynchronize(h->ctx->stream);
    for (void *p : h->owned) cudaFree(p);
    delete h;
}

plp_status plp_line_create(plp_ctx *ctx, int rows, int cols, int max_batch, plp_line **out) {
    PLP_REQUIRE(ctx && out, "null pointer");
    *out = nullptr;
    PLP_REQUIRE(rows >= 16 && cols >= 16 && rows < 32768 && cols < 32768, "image size");
    PLP_REQUIRE(max_batch >= 1, "max_batch");
    PLP_CUDA_TRY(cudaSetDevice(ctx->device));
    plp_line *h = new plp_line();
    h->ctx = ctx;
    h->rows = rows;
    h->cols = cols;
    h->max_batch = max_batch;
    LineDev &D = h->dev;
    D.w = cols;
    D.h = rows;
    D.sw = (int)lrint(cols * 0.5);
    D.sh = (int)lrint(rows * 0.5);
    D.npx = D.sw * D.sh;
    // line_extractor.cc:113-122 / lsd.cpp flsd
    const double ang_th = 22.5, quant = 2.0;
    D.prec = kPi * ang_th / 180;
    D.p = ang_th / 180;
    D.rho = quant / std::sin(D.prec);
    D.density_th = 0.6;
    const double log_nt = 5 * (std::log10((double)D.sw) + std::log10((double)D.sh)) / 2 + std::log10(11.0);
    D.min_reg_size = (int)(size_t)(-log_nt / std::log10(D.p));
    D.min_length = 0.125 * std::min(cols, rows);
    D.seg_cap = D.npx / std::max(D.min_reg_size, 2) + 1;
    {  // largest squared norm with sqrt(k / 4.0) <= rho (lsd.cpp: `norm <= threshold` -> NOTDEF)
        int k = (int)std::floor(4.0 * D.rho * D.rho) + 2;
        while (k > 0 && !(std::sqrt((double)k / 4.0) <= D.rho)) --k;
        D.kthr = k;
    }
    D.kl_cap = 1024;
    {  // LBD weights, binary_descriptor_custom.cpp:229-257 (integer divisions kept)
        double u = (kBandWidth * 3 - 1) / 2;
        double sigma = (kBandWidth * 2 + 1) / 2;
        double inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < kBandWidth * 3; ++i) {
            const double dis = i - u;
            D.gauss_l[i] = (float)std::exp(dis * dis * inv);
        }
        u = (kBands * kBandWidth - 1) / 2;
        sigma = u;
        inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < kLspHeight; ++i) {
            const double dis = i - u;
            D.gauss_g[i] = (float)std::exp(dis * dis * inv);
        }
    }
    const size_t B = max_batch, npx = D.npx;
    plp_status st = PLP_OK;
#define A(call) if (st == PLP_OK) st = (call)
    A(dev_alloc(h, &D.scaled, B * npx));
    A(dev_alloc(h, &D.order, B * npx));
    A(dev_alloc(h, &D.nseeds, B));
    A(dev_alloc(h, &D.reg_xy, B * npx));
    A(dev_alloc(h, &D.segs, B * D.seg_cap));
    A(dev_alloc(h, &D.nseg, B));
    A(dev_alloc(h, &D.grad, B * (size_t)rows * cols));
    A(dev_alloc(h, &D.lbd_float, B * D.kl_cap * 72));
    A(dev_alloc(h, &D.status, B));
    A(dev_alloc(h, &h->d_img, B * (size_t)rows * cols));
    A(dev_alloc(h, &h->d_kl, B * D.kl_cap));
    A(dev_alloc(h, &h->d_lbd, B * D.kl_cap * 32));
    A(dev_alloc(h, &h->d_fn, B * D.kl_cap * 3));
    A(dev_alloc(h, &h->d_n, B));
    if (st != PLP_OK) {
        plp_line_destroy(h);
        return st;
    }
    A(dev_alloc(h, &h->d_cstab, (size_t)kGDim * kGDim));
    if (st != PLP_OK) {
        plp_line_destroy(h);
        return st;
    }
#undef A
    D.cstab = h->d_cstab;
    lsd_cs_table_kernel<<<div_up(kGDim * kGDim, 256), 256, 0, ctx->stream>>>(h->d_cstab);
    ctx->launches++;
    h->sort_smem = ((size_t)kSortWarps * kBins + kBins) * sizeof(uint32_t);
    const size_t used_bytes = (size_t)((((D.npx + 31) >> 5) + 3) & ~3) * 4;
    D.reg_cap_small = kRegCapSmall;
    if (const char *ev = getenv("PLP_LSD_REGCAP")) D.reg_cap_small = std::max(64, std::min(kRegCap, atoi(ev)));  // tuning aid
    h->dev.reg_cap_small = D.reg_cap_small;
    h->grow_smem_noimg = used_bytes + (size_t)D.reg_cap_small * 4;
    h->grow_smem = (size_t)((D.npx + 15) & ~15) + used_bytes + (size_t)kRegCap * 4;
    h->img_smem_ok = h->grow_smem <= 227 * 1024;
    if (h->grow_smem_noimg > 227 * 1024) {
        set_error("line: a %d x %d image needs %zu bytes of shared memory per frame (limit 232448)", cols, rows,
                  h->grow_smem_noimg);
        plp_line_destroy(h);
        return PLP_ERR_CAPACITY;
    }
    h->resident_smem_frames = h->img_smem_ok ? ctx->sm_count * (int)std::max<size_t>(1, (227 * 1024) / (h->grow_smem + 1024)) : 0;
    plp_status so = PLP_OK;
    if (h->img_smem_ok) so = ensure_smem_optin((const void *)lsd_grow_kernel<true>, h->grow_smem, "lsd_grow_kernel<true>");
    if (so == PLP_OK) so = ensure_smem_optin((const void *)lsd_grow_kernel<false>, h->grow_smem_noimg, "lsd_grow_kernel<false>");
    if (so == PLP_OK) so = ensure_smem_optin((const void *)lsd_sort_kernel, h->sort_smem, "lsd_sort_kernel");
    {  // multi-warp variant: image + committed bitmap + per warp {private bitmap, region window}
        const size_t fixed = (size_t)((D.npx + 15) & ~15) + used_bytes + sizeof(MwCtl) + 64, per_warp = used_bytes + (size_t)kMwRegCap * 4;
        const size_t budget = 227 * 1024;
        h->mw_warps = fixed + 2 * per_warp <= budget ? (int)std::min<size_t>(kMwMaxWarps, (budget - fixed) / per_warp) : 0;
        if (const char *ev = getenv("PLP_LSD_DIRECT")) h->dev.direct_trig = atoi(ev);  // tuning aid
        if (const char *ev = getenv("PLP_LSD_MW_WARPS")) h->mw_warps = std::max(0, std::min(h->mw_warps, atoi(ev)));  // tuning aid
        h->mw_smem = fixed + (size_t)h->mw_warps * per_warp;
        h->mw_max_batch = h->mw_warps >= 2 ? std::min(max_batch, ctx->sm_count) : 0;
        if (h->mw_warps >= 2) {
            if (so == PLP_OK) so = ensure_smem_optin((const void *)lsd_grow_mw_kernel, h->mw_smem, "lsd_grow_mw_kernel");
            if (so == PLP_OK) so = dev_alloc(h, &h->d_reg_mw, (size_t)h->mw_max_batch * (kMwMaxWarps + 1) * D.npx);
            // out-of-order variant: one more private bitmap + window (execution at the head), the reorder buffer
            const size_t ofixed = (size_t)((D.npx + 15) & ~15) + 3 * used_bytes + (size_t)kOooRegCap * 4 + sizeof(OooEntry) * kOooRing +
                                  sizeof(OooCtl) + 64, oper = used_bytes + (size_t)kOooRegCap * 4;
            h->ooo_warps = ofixed + 2 * oper <= budget ? (int)std::min<size_t>(kMwMaxWarps, (budget - ofixed) / oper) : 0;
            if (const char *ev = getenv("PLP_LSD_OOO")) h->ooo_auto = atoi(ev) != 0;
            if (const char *ev = getenv("PLP_LSD_OOO_WARPS")) h->ooo_warps = std::max(0, std::min(h->ooo_warps, atoi(ev)));  // tuning aid
            h->ooo_smem = ofixed + (size_t)h->ooo_warps * oper;
            if (h->ooo_warps >= 2 && so == PLP_OK)
                so = ensure_smem_optin((const void *)lsd_grow_ooo_kernel, h->ooo_smem, "lsd_grow_ooo_kernel");
            if (so == PLP_OK) so = dev_alloc(h, &h->dev.mw_stat, (size_t)8 * max_batch);
            if (so == PLP_OK && cudaMemsetAsync(h->dev.mw_stat, 0, (size_t)8 * max_batch * 8, ctx->stream) != cudaSuccess) so = PLP_ERR_CUDA;
        }
    }
    if (so != PLP_OK) {
        plp_line_destroy(h);
        return so;
    }
    *out = h;
    return PLP_OK;
}

int plp_line_capacity(const plp_line *h) { return h ? h->dev.kl_cap : 0; }

plp_status plp_line_extract_batch_dev(plp_line *h, const uint8_t *d_imgs, int batch, size_t step, plp_keyline *d_kl,
                                      uint8_t *d_lbd, double *d_fn, int32_t *d_n, int32_t *d_status) {
    PLP_REQUIRE(h && d_imgs && d_kl && d_lbd && d_fn && d_n, "null pointer");
    PLP_REQUIRE(batch >= 1 && batch <= h->max_batch, "batch exceeds the handle's max_batch");
    PLP_REQUIRE(step >= (size_t)h->cols, "step < cols");
    PLP_CUDA_TRY(cudaSetDevice(h->ctx->device));
    return line_run(h, d_imgs, batch, step, d_kl, d_lbd, d_fn, d_n, d_status);
}

static plp_status line_extract_host(plp_line *h, const uint8_t *imgs, int batch, size_t step, plp_keyline *kl_out,
                                    uint8_t *lbd_out, double *fn_out, int32_t *n_out) {
    plp_ctx *ctx = h->ctx;
    PLP_CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t rows = h->rows, cols = h->cols, cap = h->dev.kl_cap;
    PLP_CUDA_TRY(cudaMemcpy2DAsync(h->d_img, cols, imgs, step, cols, rows * (size_t)batch, cudaMemcpyHostToDevice,
                                   ctx->stream));
    PLP_TRY(line_run(h, h->d_img, batch, cols, h->d_kl, h->d_lbd, h->d_fn, h->d_n, nullptr));
    PLP_CUDA_TRY(cudaMemcpyAsync(n_out, h->d_n, (size_t)batch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaMemcpyAsync(kl_out, h->d_kl, (size_t)batch * cap * sizeof(plp_keyline), cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaMemcpyAsync(lbd_out, h->d_lbd, (size_t)batch * cap * 32, cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaMemcpyAsync(fn_out, h->d_fn, (size_t)batch * cap * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<int> status(batch);
    PLP_CUDA_TRY(cudaMemcpyAsync(status.data(), h->dev.status, (size_t)batch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < batch; ++b)
        if (status[b] != 0) {
            set_error("line: capacity overflow in frame %d (code %d)", b, status[b]);
            return PLP_ERR_CAPACITY;
        }
    return PLP_OK;
}

plp_status plp_line_extract(plp_line *h, const uint8_t *img, int rows, int cols, size_t step, plp_keyline *kl_out,
                            uint8_t *lbd_out, double *fn_out, int *n_out) {
    PLP_REQUIRE(h && n_out, "null pointer");
    *n_out = 0;
    PLP_REQUIRE(img && kl_out && lbd_out && fn_out, "null pointer");
    PLP_REQUIRE(rows == h->rows && cols == h->cols, "image size differs from the handle's");
    PLP_REQUIRE(step >= (size_t)cols, "step < cols");
    int32_t n = 0;
    PLP_TRY(line_extract_host(h, img, 1, step, kl_out, lbd_out, fn_out, &n));
    *n_out = n;
    return PLP_OK;
}

plp_status plp_line_extract_batch(plp_line *h, const uint8_t *imgs, int batch, size_t step, plp_keyline *kl_out,
                                  uint8_t *lbd_out, double *fn_out, int32_t *n_out) {
    PLP_REQUIRE(h && imgs && kl_out && lbd_out && fn_out && n_out, "null pointer");
    PLP_REQUIRE(batch >= 1 && batch <= h->max_batch, "batch exceeds the handle's max_batch");
    PLP_REQUIRE(step >= (size_t)h->cols, "step < cols");
    return line_extract_host(h, imgs, batch, step, kl_out, lbd_out, fn_out, n_out);
}

plp_status plp_line_debug_force_global_image(plp_line *h, int on) {
    PLP_REQUIRE(h, "null pointer");
    h->force_global_image = on != 0;
    return PLP_OK;
}

plp_status plp_line_debug_grow_variant(plp_line *h, int variant) {
    PLP_REQUIRE(h && variant >= 0 && variant <= 3, "variant must be 0 (automatic), 1 (one warp per frame), 2 (multi-warp rounds) or 3 (out of order)");
    PLP_REQUIRE(variant < 2 || h->mw_warps >= 2, "the multi-warp variants do not fit the shared memory at this image size");
    PLP_REQUIRE(variant != 3 || h->ooo_warps >= 2, "the out-of-order variant does not fit the shared memory at this image size");
    h->grow_variant = variant;
    return PLP_OK;
}

plp_status plp_line_debug_grow_stats(plp_line *h, int b, unsigned long long *out3) {
    PLP_REQUIRE(h && out3, "null pointer");
    PLP_REQUIRE(b >= 0 && b < h->max_batch, "index");
    for (int q = 0; q < 8; ++q) out3[q] = 0;
    if (!h->dev.mw_stat) return PLP_OK;
    plp_ctx *ctx = h->ctx;
    PLP_CUDA_TRY(cudaSetDevice(ctx->device));
    PLP_CUDA_TRY(cudaMemcpyAsync(out3, h->dev.mw_stat + 8 * (size_t)b, 64, cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PLP_OK;
}

plp_status plp_line_debug_segments(plp_line *h, int b, float *segs_out, int cap, int *n_out) {
    PLP_REQUIRE(h && segs_out && n_out, "null pointer");
    PLP_REQUIRE(b >= 0 && b < h->last_batch, "index");
    plp_ctx *ctx = h->ctx;
    PLP_CUDA_TRY(cudaSetDevice(ctx->device));
    int n = 0;
    PLP_CUDA_TRY(cudaMemcpyAsync(&n, h->dev.nseg + b, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    PLP_REQUIRE(n <= cap, "cap too small");
    PLP_CUDA_TRY(cudaMemcpyAsync(segs_out, h->dev.segs + (size_t)b * h->dev.seg_cap, (size_t)n * 16, cudaMemcpyDeviceToHost,
                                 ctx->stream));
    PLP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PLP_OK;
}

plp_status plp_line_debug_scaled(plp_line *h, int b, uint8_t *out) {
    PLP_REQUIRE(h && out, "null pointer");
    PLP_REQUIRE(b >= 0 && b < h->last_batch, "index");
    plp_ctx *ctx = h->ctx;
    PLP_CUDA_TRY(cudaSetDevice(ctx->device));
    PLP_CUDA_TRY(cudaMemcpyAsync(out, h->dev.scaled + (size_t)b * h->dev.npx, h->dev.npx, cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PLP_OK;
}

plp_status plp_line_debug_lbd_float(plp_line *h, int b, float *out, int cap) {
    PLP_REQUIRE(h && out, "null pointer");
    PLP_REQUIRE(b >= 0 && b < h->last_batch && cap >= 0 && cap <= h->dev.kl_cap, "index");
    plp_ctx *ctx = h->ctx;
    PLP_CUDA_TRY(cudaSetDevice(ctx->device));
    PLP_CUDA_TRY(cudaMemcpyAsync(out, h->dev.lbd_float + (size_t)b * h->dev.kl_cap * 72, (size_t)cap * 72 * sizeof(float),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    PLP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PLP_OK;
}

}  // extern "C"
