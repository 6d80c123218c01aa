// oracle/frontend_mt.cc -- the per-frame front-end chain of the oracle on a native thread pool
// (TEST / BENCH INFRASTRUCTURE ONLY: bench.py's cpu_baseline and --impl reference legs).
//
// One frame = what tracking_module::track does with frame_tracker::motion_based_track
// (module/frame_tracker.cc:52-124): orb_extractor::extract (feature/orb_extractor.cc:73-160) ->
// projection::match_current_and_last_frames with `margin`, retried with 2 * margin below 20 matches
// (frame_tracker.cc:63-71) -> pose_optimizer::optimize (optimize/pose_optimizer.cc:53-229) -> discard_outliers
// (frame_tracker.cc:253-283).  Frames are independent tracking problems, so they are spread over `threads`
// std::threads (work stealing by an atomic counter, each thread optionally pinned to one CPU), all inside this
// library: no interpreter, no GIL hand-offs, so two runs on the same box agree.  The per-frame functions are the
// single-threaded restatements the parity tests use.
#include <sched.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "oracle.h"

namespace {
// the CPUs this process may run on (cgroup / taskset aware); thread t is pinned to allowed[t % size]
std::vector<int> allowed_cpus() {
    std::vector<int> v;
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0)
        for (int c = 0; c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &set)) v.push_back(c);
    return v;
}
void pin_thread(const std::vector<int> &cpus, int tix) {
    if (cpus.empty()) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(cpus[tix % cpus.size()], &set);
    sched_setaffinity(0, sizeof(set), &set);
}
}  // namespace

extern "C" {

/* number of CPUs the calling process may be scheduled on */
int orc_host_cpus(void) { return (int)allowed_cpus().size(); }

/* returns the number of frames processed; seconds_out = wall time of the parallel region (steady_clock) */
int orc_frontend_track_batch(const orc_orb_params *p, const orc_grid *g, const orc_camera *cam, const uint8_t *imgs, int batch,
                             int rows, int cols, const double *last_pos_w, const int32_t *last_octave,
                             const float *last_angle, const uint8_t *last_desc, const uint8_t *last_valid,
                             const int32_t *last_offsets, const double *pose_pred, const double *pose_last, float margin,
                             int threads, int pin, double *pose_out, int32_t *n_inliers_out, int32_t *n_valid_out,
                             int32_t *n_kp_out, double *seconds_out) {
    if (threads < 1) threads = 1;
    const int L = (int)p->num_levels;
    std::vector<float> sf(L), isf(L), ls(L), ils(L);
    std::vector<uint32_t> nk(L);
    int32_t umax[16];
    orc_orb_tables(p, sf.data(), isf.data(), ls.data(), ils.data(), nk.data(), umax);
    const orc_pose_cam pc{cam->fx, cam->fy, cam->cx, cam->cy, cam->focal_x_baseline, cam->setup_type};
    const int cap = 4 * (int)p->max_num_keypts + 64 * L;
    std::atomic<int> next(0);
    const std::vector<int> cpus = allowed_cpus();
    auto worker = [&](int tix) {
        if (pin && threads > 1) pin_thread(cpus, tix);
        std::vector<orc_keypoint> kps(cap);
        std::vector<uint8_t> desc((size_t)cap * 32);
        std::vector<float> x(cap), y(cap), ang(cap);
        std::vector<int32_t> oct(cap), matched(cap), obs_kp(cap);
        std::vector<orc_pt_obs> obs(cap);
        std::vector<uint8_t> outl(cap);
        for (;;) {
            const int b = next.fetch_add(1);
            if (b >= batch) break;
            const int n = orc_orb_extract(p, imgs + (size_t)b * rows * cols, rows, cols, cols, nullptr, 0, kps.data(),
                                          desc.data(), cap, nullptr, nullptr, 0, nullptr, nullptr);
            n_kp_out[b] = n;
            n_inliers_out[b] = 0;
            n_valid_out[b] = 0;
            memcpy(pose_out + 16 * (size_t)b, pose_pred + 16 * (size_t)b, 128);
            if (n <= 0) continue;
            for (int i = 0; i < n; ++i) {
                x[i] = kps[i].x;
                y[i] = kps[i].y;
                ang[i] = kps[i].angle;
                oct[i] = kps[i].octave;
            }
            const int l0 = last_offsets[b], m = last_offsets[b + 1] - l0;
            unsigned nm = 0;
            for (int attempt = 0; attempt < 2; ++attempt) {  // frame_tracker.cc:63-71
                nm = orc_match_current_and_last_frames(g, n, x.data(), y.data(), oct.data(), ang.data(), nullptr, desc.data(),
                                                       nullptr, sf.data(), L, cam, pose_pred + 16 * (size_t)b,
                                                       pose_last + 16 * (size_t)b, m, last_pos_w + 3 * (size_t)l0,
                                                       last_octave + l0, last_angle + l0, last_desc + 32 * (size_t)l0,
                                                       last_valid ? last_valid + l0 : nullptr, (attempt ? 2.0f : 1.0f) * margin,
                                                       1, matched.data());
                if (nm >= 20) break;
            }
            if (nm < 20) continue;  // frame_tracker.cc:73-77
            int no = 0;
            for (int i = 0; i < n; ++i) {  // pose_optimizer.cc:126-151, keypoint order
                const int q = matched[i];
                if (q < 0) continue;
                orc_pt_obs &o = obs[no];
                const double *X = last_pos_w + 3 * (size_t)(l0 + q);
                o.pos_w[0] = X[0];
                o.pos_w[1] = X[1];
                o.pos_w[2] = X[2];
                o.obs_x = x[i];
                o.obs_y = y[i];
                o.x_right = -1.0f;
                o.inv_sigma_sq = ils[oct[i]];
                obs_kp[no++] = i;
            }
            int iters = 0;
            n_inliers_out[b] = orc_pose_optimize(&pc, pose_pred + 16 * (size_t)b, obs.data(), no, nullptr, 0, 4, 10,
                                                 pose_out + 16 * (size_t)b, outl.data(), nullptr, &iters);
            int valid = 0;
            for (int k = 0; k < no; ++k) valid += outl[k] ? 0 : 1;  // discard_outliers (frame_tracker.cc:253-283)
            n_valid_out[b] = valid;
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
        for (auto &t : pool) t.join();
    }
    if (seconds_out) *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return batch;
}

/* LSD + LBD extraction (feature/line_extractor.cc:88-160) of `batch` frames on `threads` native threads */
int orc_line_extract_batch_mt(const uint8_t *imgs, int batch, int rows, int cols, const orc_lsd_config *cfg, int threads, int pin,
                              int32_t *n_out, double *seconds_out) {
    if (threads < 1) threads = 1;
    std::atomic<int> next(0);
    const std::vector<int> cpus = allowed_cpus();
    const int cap = 8192;
    auto worker = [&](int tix) {
        if (pin && threads > 1) pin_thread(cpus, tix);
        std::vector<orc_keyline> kl(cap);
        std::vector<uint8_t> lbd((size_t)cap * 32);
        std::vector<double> fn((size_t)cap * 3);
        for (;;) {
            const int b = next.fetch_add(1);
            if (b >= batch) break;
            n_out[b] = orc_line_extract(imgs + (size_t)b * rows * cols, cols, rows, cols, cfg, kl.data(), lbd.data(), fn.data(), cap);
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
        for (auto &t : pool) t.join();
    }
    if (seconds_out) *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return batch;
}

/* one rectified stereo frame of frame::frame (data/frame.cc:456-480): ORB left + right, match::stereo::compute, LSD + LBD
 * left + right; frames spread over `threads` native threads */
int orc_stereo_frontend_batch_mt(const orc_orb_params *p, const orc_lsd_config *cfg, const uint8_t *left, const uint8_t *right,
                                 int batch, int rows, int cols, float focal_x_baseline, float true_baseline, int threads,
                                 int pin, int32_t *n_kp_out, int32_t *n_stereo_out, int32_t *n_lines_out, double *seconds_out) {
    if (threads < 1) threads = 1;
    const int L = (int)p->num_levels;
    std::vector<float> sf(L), isf(L), ls(L), ils(L);
    std::vector<uint32_t> nk(L);
    int32_t umax[16];
    orc_orb_tables(p, sf.data(), isf.data(), ls.data(), ils.data(), nk.data(), umax);
    std::vector<int32_t> lw(L), lh(L);
    orc_orb_level_sizes(p, rows, cols, lw.data(), lh.data());
    size_t pyr_bytes = 0;
    for (int l = 0; l < L; ++l) pyr_bytes += (size_t)lw[l] * lh[l];
    const int cap = 4 * (int)p->max_num_keypts + 64 * L, lcap = 8192;
    std::atomic<int> next(0);
    const std::vector<int> cpus = allowed_cpus();
    auto worker = [&](int tix) {
        if (pin && threads > 1) pin_thread(cpus, tix);
        std::vector<orc_keypoint> kl_(cap), kr_(cap);
        std::vector<uint8_t> dl((size_t)cap * 32), dr((size_t)cap * 32), pl(pyr_bytes), pr(pyr_bytes);
        std::vector<float> xr(cap), dp(cap);
        std::vector<orc_keyline> kl(lcap);
        std::vector<uint8_t> lbd((size_t)lcap * 32);
        std::vector<double> fn((size_t)lcap * 3);
        for (;;) {
            const int b = next.fetch_add(1);
            if (b >= batch) break;
            const uint8_t *il = left + (size_t)b * rows * cols, *ir = right + (size_t)b * rows * cols;
            const int nl = orc_orb_extract(p, il, rows, cols, cols, nullptr, 0, kl_.data(), dl.data(), cap, pl.data(), nullptr,
                                           0, nullptr, nullptr);
            const int nr = orc_orb_extract(p, ir, rows, cols, cols, nullptr, 0, kr_.data(), dr.data(), cap, pr.data(), nullptr,
                                           0, nullptr, nullptr);
            n_kp_out[2 * b] = nl;
            n_kp_out[2 * b + 1] = nr;
            int ns = 0;
            if (nl > 0 && nr > 0) {
                orc_stereo_compute(pl.data(), pr.data(), lw.data(), lh.data(), L, kl_.data(), dl.data(), nl, kr_.data(),
                                   dr.data(), nr, sf.data(), isf.data(), focal_x_baseline, true_baseline, xr.data(), dp.data(),
                                   nullptr);
                for (int i = 0; i < nl; ++i) ns += dp[i] > 0.0f ? 1 : 0;
            }
            n_stereo_out[b] = ns;
            n_lines_out[2 * b] = orc_line_extract(il, cols, rows, cols, cfg, kl.data(), lbd.data(), fn.data(), lcap);
            n_lines_out[2 * b + 1] = orc_line_extract(ir, cols, rows, cols, cfg, kl.data(), lbd.data(), fn.data(), lcap);
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
        for (auto &t : pool) t.join();
    }
    if (seconds_out) *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return batch;
}

}  // extern "C"
