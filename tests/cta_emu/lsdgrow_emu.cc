// lsdgrow_emu.cc -- csrc/lsd_grow_kernels.cuh (LSD region growing: one warp per frame, the multi-warp round protocol, the
// out-of-order variant with a reorder buffer) executed on the host, one host thread per CUDA thread.  The warps of the
// multi-warp variants really run concurrently here (host threads), so their locks, tickets and in-order commit are exercised.
#include "cta_emu.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "lsd_grow_kernels.cuh"

using namespace plp::lsd;

// variant 1: one warp, 2: rounds, 3: out of order.  order_xy: seeds (y << 16 | x) in processing order.
// Returns 0, or -1 when the variant / warp count is not available.
extern "C" int emu_lsd_grow(int variant, int warps, const uint8_t *scaled, int sw, int sh, const uint32_t *order_xy, int nseeds,
                            float *segs_out, int seg_cap, int *nseg_out, int *status_out, unsigned long long *stat_out) {
    LineDev D;
    memset(&D, 0, sizeof(D));
    D.sw = sw;
    D.sh = sh;
    D.npx = sw * sh;
    D.w = 2 * sw;
    D.h = 2 * sh;
    const double ang_th = 22.5, quant = 2.0;  // lines.cu plp_line_create (line_extractor.cc:113-122 / lsd.cpp flsd)
    D.prec = kPi * ang_th / 180;
    D.p = ang_th / 180;
    D.rho = quant / sin(D.prec);
    D.density_th = 0.6;
    const double log_nt = 5 * (log10((double)sw) + log10((double)sh)) / 2 + log10(11.0);
    D.min_reg_size = (int)(size_t)(-log_nt / log10(D.p));
    D.seg_cap = seg_cap;
    {
        int k = (int)floor(4.0 * D.rho * D.rho) + 2;
        while (k > 0 && !(sqrt((double)k / 4.0) <= D.rho)) --k;
        D.kthr = k;
    }
    static std::vector<float4> tab;  // {deg, cos, sin} by (gx, gy), as lsd_cs_table_kernel builds it
    if (tab.empty()) {
        tab.resize((size_t)kGDim * kGDim);
        for (int i = 0; i < kGDim * kGDim; ++i) {
            const int gy = i / kGDim - kGRange, gx = i - (gy + kGRange) * kGDim - kGRange;
            const float deg = fast_atan2_deg((float)gx, (float)-gy);
            const double a = (double)deg * kDegToRads;
            const float af = (float)a;
            tab[i] = make_float4(deg, (float)det_cos((double)af), (float)det_sin((double)af), 0.f);
        }
    }
    std::vector<uint8_t> img(scaled, scaled + D.npx);
    std::vector<uint32_t> order(order_xy, order_xy + nseeds), reg_xy((size_t)D.npx), ovf((size_t)(kMwMaxWarps + 1) * D.npx);
    std::vector<float4> segs((size_t)seg_cap);
    int nseeds_v = nseeds, nseg = 0, status = 0;
    unsigned long long stat[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    D.scaled = img.data();
    D.cstab = tab.data();
    D.order = order.data();
    D.nseeds = &nseeds_v;
    D.reg_xy = reg_xy.data();
    D.segs = segs.data();
    D.nseg = &nseg;
    D.status = &status;
    D.mw_stat = stat;
    D.reg_cap_small = kRegCapSmall;
    const size_t smem = 1u << 20;
    if (variant == 1) {
        emu_launch2(lsd_grow_kernel<true>, 1u, 1u, 32u, smem, D);
    } else if (variant == 2 && warps >= 2 && warps <= kMwMaxWarps) {
        emu_launch2(lsd_grow_mw_kernel, 1u, 1u, (unsigned)warps * 32u, smem, D, ovf.data());
    } else if (variant == 3 && warps >= 2 && warps <= kMwMaxWarps) {
        emu_launch2(lsd_grow_ooo_kernel, 1u, 1u, (unsigned)warps * 32u, smem, D, ovf.data());
    } else {
        return -1;
    }
    *nseg_out = nseg;
    *status_out = status;
    for (int i = 0; i < std::min(nseg, seg_cap); ++i) {
        segs_out[4 * i] = segs[i].x;
        segs_out[4 * i + 1] = segs[i].y;
        segs_out[4 * i + 2] = segs[i].z;
        segs_out[4 * i + 3] = segs[i].w;
    }
    memcpy(stat_out, stat, sizeof(stat));
    return 0;
}
