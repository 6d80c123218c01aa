#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Structure-PLP-SLAM hot path.

Metric (BASELINE.json): frames/s through extract + match + pose-opt on synthetic 640x480 sequences
(8-level ORB pyramid, ~1000 kp/frame).  A "step" = one pass of the front-end
    orb_extractor::extract -> frame_tracker::motion_based_track
      (match_current_and_last_frames [+ widened retry] -> pose_optimizer::optimize -> discard_outliers)
over one batch of independent frames (frames shard over GPUs with no collective).

  value     : device-resident throughput (images already in HBM when the timed region starts)
  e2e       : same metric through the public batched API with HOST buffers: pinned H2D of the step's images and
              D2H of the step's poses / inlier counts inside the timed region
  roofline  : dominant kernel's algorithmic bytes per launch / its mean launch time (CUDA events on the
              launching stream) vs the measured HBM peak
  cpu_baseline / --impl reference : the CPU oracle port of the same path on the host cores (the reference's
              own binary cannot be built here: no C++ OpenCV / Eigen / g2o), bounded sample.

Launch:  python bench.py --gpus N --steps K --warmup W          (N>1: via torch.distributed.run, one rank per GPU)
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "frames_per_sec_extract_match_poseopt"
UNIT = "frames/s"
ROWS, COLS = 480, 640
WORKLOAD = "ORB extract + Hamming match + pose-opt, 640x480, 8-level pyramid, 1000 kp/frame (BASELINE configs[1]+[2])"
# algorithmic bytes per frame of the extraction path; per-kernel split in DESIGN.md
PYR_PX = 950532
ALG_BYTES = {
    "fast_cells_kernel": PYR_PX,  # every pyramid level read once for FAST (+ candidates out, negligible)
    "fast_cells_kernel_v2": PYR_PX,
    "pyr_resize_kernel": None,    # per level, filled below
    "blur_tiles_kernel": 2 * PYR_PX,          # pyramid read + blurred pyramid written
    "describe_kernel": PYR_PX + 60 * 1200,
    "quadtree_kernel": 8 * 11000 + 8 * 1200,
    # match (per frame) = (M + N) x 48 B: 32 B descriptor + 16 B geometry on each side, N ~ 1200 keypoints, M ~ 1000 queries
    "point_match_kernel": 48 * (1200 + 1000),
    # pose-opt = 41 B per point edge and LM iteration (24 B position + 12 B observation + 4 B weight + 1 B flag), <= 40
    # iterations + 4 re-classification passes per call; ~1000 edges per frame
    "pose_opt_kernel": 41 * 1000 * 44,
}


def _load_pkg():
    if "plpslam_b200" in sys.modules:
        return sys.modules["plpslam_b200"]
    pkg = ROOT / "structure-plp-slam_b200"
    spec = importlib.util.spec_from_file_location("plpslam_b200", pkg / "__init__.py",
                                                  submodule_search_locations=[str(pkg)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["plpslam_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, device: int):
        self.device = device
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        # NVML in-process (a sample every 10 ms, so even a 100 ms timed region yields several samples under load);
        # nvidia-smi as the fallback (one sample per ~150 ms)
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [(0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap")]
            while not self._stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                self.samples.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b, _ in bits])
                self._stop.wait(0.01)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max(int(s[1]) for s in self.samples if s[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.samples)}


def build_inputs(batch: int, seed: int):
    """`batch` independent tracking problems from a few rendered planar sequences (tests/scene.py)."""
    import oracle_api
    import scene
    n_seq_frames = 33
    seqs = []
    frames, t_idx = [], []
    s = 0
    while len(frames) < batch:
        seq = scene.PlanarSequence(seed=seed + 17 * s, n_frames=n_seq_frames)
        seqs.append(seq)
        for t in range(1, n_seq_frames):
            if len(frames) < batch:
                frames.append(seq.frames[t])
                t_idx.append((s, t))
        s += 1
    return seqs, np.stack(frames), t_idx


def setup_front_end(pkg, ctx, batch, seed, track_ctx=None):
    """Render frames, extract the 'last' frames once (untimed) to obtain landmark sets, upload everything."""
    import synth
    from plpslam_b200.tracking import FrontEnd
    cam = pkg.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, COLS, ROWS)
    fe = FrontEnd(ctx, ROWS, COLS, cam, max_batch=batch, track_ctx=track_ctx)
    seqs, frames, t_idx = build_inputs(batch, seed)
    # last-frame landmarks: extract frame t-1 of every problem on the GPU (setup, untimed)
    last_imgs = np.stack([seqs[s].frames[t - 1] for (s, t) in t_idx])
    fe.upload_images(last_imgs)
    fe.extract(batch)
    ctx.sync()
    kps = fe.download_keypoints(batch)
    rng = np.random.default_rng(seed)
    lasts = [seqs[s].last_frame_landmarks(t - 1, kps[b][0], kps[b][1]) for b, (s, t) in enumerate(t_idx)]
    preds = np.stack([seqs[s].predicted_pose(t, rng) for (s, t) in t_idx])
    pose_last = np.stack([seqs[s].poses[t - 1] for (s, t) in t_idx])
    fe.set_last_frames(lasts, preds, pose_last)
    fe.upload_images(frames)
    gt = np.stack([seqs[s].poses[t] for (s, t) in t_idx])
    return fe, frames, dict(seqs=seqs, t_idx=t_idx, lasts=lasts, preds=preds, gt=gt)


def cpu_port_frames(frames, aux, idxs, threads):
    """The oracle port of the same path (extract -> match -> pose-opt) on `threads` host threads."""
    import oracle_api
    import synth
    pkg = _load_pkg()
    orc = oracle_api.Oracle()
    p = oracle_api.orb_params()
    grid = pkg.capi.make_grid(COLS, ROWS)
    cam = pkg.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, COLS, ROWS)
    sf, isig = synth.scale_factors(), synth.inv_level_sigma_sq()

    def one(b):
        r = orc.orb_extract(p, frames[b])
        k = r["kps"]
        last = aux["lasts"][b]
        s, t = aux["t_idx"][b]
        curr = dict(x=k["x"], y=k["y"], octave=k["octave"], angle=k["angle"], desc=r["desc"])
        Tp = aux["preds"][b]
        Tl = aux["seqs"][s].poses[t - 1]
        m, nm = orc.match_current_and_last_frames(grid, sf, cam, curr, Tp, Tl, last, 20.0, True)
        if nm < 20:
            m, nm = orc.match_current_and_last_frames(grid, sf, cam, curr, Tp, Tl, last, 40.0, True)
        idx = np.nonzero(m >= 0)[0]
        pts = np.zeros(len(idx), oracle_api.PT_OBS_DTYPE)
        pts["pos_w"] = last["pos_w"][m[idx]]
        pts["obs_x"], pts["obs_y"] = k["x"][idx], k["y"][idx]
        pts["x_right"] = -1.0
        pts["inv_sigma_sq"] = isig[k["octave"][idx]]
        T, pout, _, n_inl, _ = orc.pose_optimize(cam, Tp, pts)
        return n_inl

    t0 = time.perf_counter()
    if threads <= 1:
        out = [one(b) for b in idxs]
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            out = list(ex.map(one, idxs))
    dt = time.perf_counter() - t0
    return len(idxs) / dt, dt, out


def build_line_frames(batch: int, seed: int):
    """`batch` distinct 640x480 line-rich frames: a few rendered scenes (tests/synth.make_line_image), each shifted by a
    different offset so that no two frames of a batch are equal."""
    import synth
    n_base = min(batch, 12)
    base = [synth.make_line_image(seed + i, ROWS, COLS) for i in range(n_base)]
    rng = np.random.default_rng(seed)
    out = np.empty((batch, ROWS, COLS), np.uint8)
    for b in range(batch):
        dx, dy = (0, 0) if b < n_base else (int(rng.integers(-60, 61)), int(rng.integers(-40, 41)))
        out[b] = np.roll(base[b % n_base], (dy, dx), axis=(0, 1))
    return out


# remap is skipped (identity); per frame the line front end must read the image twice (LSD scale pass and
# LBD blur+Sobel pass), write + read the half-resolution image, write the per-pixel level-line record once (16 B at
# quarter resolution) and the int16 gradient pair once (4 B per pixel), and read both back at least once.
def line_alg_bytes(rows, cols):
    px, spx = rows * cols, (rows // 2) * (cols // 2)
    return 2 * px + 2 * spx + 2 * 16 * spx + 2 * 4 * px


def bench_lines(pkg, ctx, stream, rank, world, steps, warmup, batch, seed, cpu_baseline):
    """LSD + LBD extraction (line_extractor.cc:88-160) of `batch` frames per step, device-resident; frames shard over
    ranks with no collective."""
    import torch
    import torch.distributed as dist
    from plpslam_b200.tracking import DeviceBuffer
    lib = pkg.lib()
    frames = build_line_frames(batch, seed + 1000 * rank)
    trk = pkg.LineFeatureTracker(ctx, ROWS, COLS, max_batch=batch)
    cap = trk.capacity
    d_imgs = DeviceBuffer.from_array(ctx, frames)
    d_kl = DeviceBuffer(ctx, batch * cap * pkg.KEYLINE_DTYPE.itemsize)
    d_lbd = DeviceBuffer(ctx, batch * cap * 32)
    d_fn = DeviceBuffer(ctx, batch * cap * 24)
    d_n = DeviceBuffer(ctx, batch * 4)
    d_st = DeviceBuffer(ctx, batch * 4)

    def step():
        ctx._check(lib.plp_line_extract_batch_dev(trk.handle, d_imgs.ptr, C.c_int(batch), C.c_size_t(COLS), d_kl.ptr,
                                                  d_lbd.ptr, d_fn.ptr, d_n.ptr, d_st.ptr))

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(warmup, 3)):
        step()
    barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    launches = ctx.launch_count() - l0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=torch.cuda.current_device())
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    n = d_n.download(np.int32, (batch,))
    st = d_st.download(np.int32, (batch,))
    assert not st.any(), "line extraction capacity overflow"
    # e2e: pinned host frames in, keylines / descriptors / line functions out, through the host entry point
    kl = np.zeros((batch, cap), pkg.KEYLINE_DTYPE)
    lbd = np.zeros((batch, cap, 32), np.uint8)
    fn = np.zeros((batch, cap, 3), np.float64)
    nn = np.zeros(batch, np.int32)

    def e2e_step():
        ctx._check(lib.plp_line_extract_batch(trk.handle, frames.ctypes.data_as(C.c_void_p), C.c_int(batch),
                                              C.c_size_t(COLS), kl.ctypes.data_as(C.c_void_p), lbd.ctypes.data_as(C.c_void_p),
                                              fn.ctypes.data_as(C.c_void_p), nn.ctypes.data_as(C.c_void_p)))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(2, steps // 3)):
        e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / max(2, steps // 3)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=torch.cuda.current_device())
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    # per-kernel shares
    ctx._check(lib.plp_ctx_kernel_timing(ctx.handle, 1))
    for _ in range(min(steps, 3)):
        step()
    buf = C.create_string_buffer(1 << 16)
    ctx._check(lib.plp_ctx_kernel_timing_report(ctx.handle, buf, C.c_size_t(len(buf))))
    ctx._check(lib.plp_ctx_kernel_timing(ctx.handle, 0))
    kt = json.loads(buf.value.decode())
    tot = sum(v["total_ms"] for v in kt.values())
    shares = {k: round(v["total_ms"] / tot, 4) for k, v in sorted(kt.items(), key=lambda kv: -kv[1]["total_ms"])}
    per_launch = {k: round(v["total_ms"] / v["count"], 4) for k, v in kt.items()}
    peak, _ = _peaks()
    alg = line_alg_bytes(ROWS, COLS) * batch
    res = {"metric": "frames_per_sec_lsd_lbd_extract", "value": world * batch * steps / (ms * 1e-3), "unit": "frames/s",
           "ms_per_step": ms / steps, "scaling": "weak",
           "config": {"workload": "LSD (refine 1, scale 0.5) + LBD line extraction, 640x480, lines >= 60 px kept",
                      "frames_per_step_per_gpu": batch, "mean_keylines_per_frame": float(n.mean()),
                      "l2": "per-step working set (3.8 MB/frame of intermediates) larger than L2"},
           "e2e": {"value": world * batch / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(frames.nbytes),
                   "d2h_bytes_per_step": int(kl.nbytes + lbd.nbytes + fn.nbytes + nn.nbytes)},
           "gpu_launches": int(launches), "kernel_time_shares": shares, "ms_per_launch": per_launch,
           "algorithmic_bytes_per_step": alg, "hbm_roofline_frac": alg / (ms / steps * 1e-3) / 1e9 / peak}
    if cpu_baseline and rank == 0:
        import oracle_api
        orc = oracle_api.Oracle()
        cores = os.cpu_count() or 1
        ns = int(min(batch, max(32, 2 * cores)))
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            list(ex.map(lambda b: orc.line_extract(frames[b]), range(ns)))
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        for b in range(8):
            orc.line_extract(frames[b])
        d1 = time.perf_counter() - t1
        res["cpu_baseline"] = {"value": ns / dt, "unit": "frames/s", "cores": cores, "kind": "port",
                               "sample": f"{ns} frames over {cores} host threads ({dt:.2f} s)",
                               "single_thread_value": 8 / d1}
    trk.close()
    for d in (d_imgs, d_kl, d_lbd, d_fn, d_n, d_st):
        d.free()
    return res


def bench_stereo(pkg, ctx, stream, rank, world, steps, warmup, batch, seed):
    """BASELINE configs[4]: EuRoC-like rectified stereo 752x480, full point + line front end per stereo frame, frames
    sharded over ranks with no collective: ORB left + ORB right (frame.cc:456-457), match::stereo::compute
    (frame.cc:470-480), LSD + LBD left + right (frame.cc:458-463)."""
    import torch
    import torch.distributed as dist
    import synth
    from plpslam_b200.tracking import DeviceBuffer
    lib = pkg.lib()
    H, W = 480, 752
    n_base = min(batch, 6)
    pairs = [synth.make_stereo_pair(seed + 31 * i + 1000 * rank, H, W) for i in range(n_base)]
    rng = np.random.default_rng(seed)
    left = np.empty((batch, H, W), np.uint8)
    right = np.empty((batch, H, W), np.uint8)
    for b in range(batch):
        sh = (0, 0) if b < n_base else (int(rng.integers(-40, 41)), int(rng.integers(-60, 61)))
        left[b] = np.roll(pairs[b % n_base][0], sh, axis=(0, 1))
        right[b] = np.roll(pairs[b % n_base][1], sh, axis=(0, 1))
    # two contexts = two streams, like the reference's left / right extraction threads (frame.cc:456-463): the right
    # image's ORB + line extraction runs on the second stream; match::stereo::compute waits for both ORB passes
    ctx_r = pkg.Context(ctx.device)
    el = pkg.OrbExtractor(ctx, H, W, max_batch=batch)
    er = pkg.OrbExtractor(ctx_r, H, W, max_batch=batch)
    ll = pkg.LineFeatureTracker(ctx, H, W, max_batch=batch)
    lr = pkg.LineFeatureTracker(ctx_r, H, W, max_batch=batch)
    cap, lcap = el.capacity, ll.capacity
    d_l, d_r = DeviceBuffer.from_array(ctx, left), DeviceBuffer.from_array(ctx, right)
    kp = [DeviceBuffer(ctx, batch * cap * pkg.KP_DTYPE.itemsize) for _ in range(2)]
    ds = [DeviceBuffer(ctx, batch * cap * 32) for _ in range(2)]
    nk = [DeviceBuffer(ctx, batch * 4) for _ in range(2)]
    st = [DeviceBuffer(ctx, batch * 4) for _ in range(4)]
    d_xr, d_dp = DeviceBuffer(ctx, batch * cap * 4), DeviceBuffer(ctx, batch * cap * 4)
    kl = [DeviceBuffer(ctx, batch * lcap * pkg.KEYLINE_DTYPE.itemsize) for _ in range(2)]
    lb = [DeviceBuffer(ctx, batch * lcap * 32) for _ in range(2)]
    fn = [DeviceBuffer(ctx, batch * lcap * 24) for _ in range(2)]
    nl = [DeviceBuffer(ctx, batch * 4) for _ in range(2)]
    bf, baseline = 47.906, 0.11

    def step():
        ctx._check(lib.plp_orb_extract_batch_dev(er.handle, d_r.ptr, C.c_int(batch), C.c_size_t(W), kp[1].ptr, ds[1].ptr,
                                                 nk[1].ptr, st[1].ptr))
        ctx._check(lib.plp_orb_extract_batch_dev(el.handle, d_l.ptr, C.c_int(batch), C.c_size_t(W), kp[0].ptr, ds[0].ptr,
                                                 nk[0].ptr, st[0].ptr))
        ctx._check(lib.plp_ctx_wait_ctx(ctx.handle, ctx_r.handle))   # the right keypoints / pyramid are ready
        ctx._check(lib.plp_line_extract_batch_dev(lr.handle, d_r.ptr, C.c_int(batch), C.c_size_t(W), kl[1].ptr, lb[1].ptr,
                                                  fn[1].ptr, nl[1].ptr, st[3].ptr))
        ctx._check(lib.plp_stereo_compute_batch_dev(ctx.handle, el.handle, er.handle, C.c_int(batch), kp[0].ptr, ds[0].ptr,
                                                    nk[0].ptr, kp[1].ptr, ds[1].ptr, nk[1].ptr, C.c_float(bf),
                                                    C.c_float(baseline), d_xr.ptr, d_dp.ptr, None))
        ctx._check(lib.plp_line_extract_batch_dev(ll.handle, d_l.ptr, C.c_int(batch), C.c_size_t(W), kl[0].ptr, lb[0].ptr,
                                                  fn[0].ptr, nl[0].ptr, st[2].ptr))
        # the next step's right-image ORB pass overwrites the right pyramid: it must wait for this stereo match
        ctx._check(lib.plp_ctx_wait_ctx(ctx_r.handle, ctx.handle))

    def barrier():
        ctx.sync()
        ctx_r.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(warmup, 3)):
        step()
    barrier()
    l0 = ctx.launch_count() + ctx_r.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    ctx._check(lib.plp_ctx_wait_ctx(ctx.handle, ctx_r.handle))
    e1.record(stream)
    barrier()
    launches = ctx.launch_count() + ctx_r.launch_count() - l0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=torch.cuda.current_device())
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    xr = d_xr.download(np.float32, (batch, cap))
    n_left = nk[0].download(np.int32, (batch,))
    n_lines = nl[0].download(np.int32, (batch,))
    stereo_ok = float(np.mean([(xr[b, :n_left[b]] >= 0).sum() for b in range(batch)]))
    ctx._check(lib.plp_ctx_kernel_timing(ctx.handle, 1))
    for _ in range(min(steps, 2)):
        step()
    buf = C.create_string_buffer(1 << 16)
    ctx._check(lib.plp_ctx_kernel_timing_report(ctx.handle, buf, C.c_size_t(len(buf))))
    ctx._check(lib.plp_ctx_kernel_timing(ctx.handle, 0))
    kt = json.loads(buf.value.decode())
    tot = sum(v["total_ms"] for v in kt.values())
    shares = {k: round(v["total_ms"] / tot, 4) for k, v in sorted(kt.items(), key=lambda kv: -kv[1]["total_ms"])}
    for o in (el, er, ll, lr):
        o.close()
    ctx_r.close()
    for d in [d_l, d_r, d_xr, d_dp] + kp + ds + nk + st + kl + lb + fn + nl:
        d.free()
    return {"metric": "stereo_frames_per_sec_point_line_frontend", "value": world * batch * steps / (ms * 1e-3),
            "unit": "stereo frames/s", "ms_per_step": ms / steps, "scaling": "weak",
            "config": {"workload": "rectified stereo 752x480: ORB L+R, match::stereo::compute, LSD+LBD L+R (BASELINE configs[4])",
                       "stereo_frames_per_step_per_gpu": batch, "mean_left_keypoints": float(n_left.mean()),
                       "mean_stereo_matches": stereo_ok, "mean_left_keylines": float(n_lines.mean()),
                       "parallelism": f"stereo frames sharded over {world} GPU(s), no data-path collective"},
            "gpu_launches": int(launches), "kernel_time_shares": shares}


def bench_ba(pkg, ctx, stream, rank, world, steps, warmup):
    """Second BASELINE metric: local-BA LM iterations/s on config 4 (20 local + 10 fixed KF, 4000 points + 800 lines
    + 200 plane-owned points, ~29 k edges), landmark-sharded over `world` GPUs with one NCCL all-reduce of the packed
    reduced camera system per LM try (strong scaling: the problem size is fixed)."""
    import torch
    import torch.distributed as dist
    import ba_data
    from plpslam_b200.ba import BaComm, LocalBA, shard_boundaries, shard_edges
    prob = ba_data.make_ba_problem(42)
    comm = None
    if world > 1:
        uid = [BaComm.unique_id(ctx) if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = BaComm(ctx, uid[0], world, rank)
        sub = prob.shard(world, rank, shard_boundaries, shard_edges)
    else:
        sub = prob
    st = sub.struct()
    ba = LocalBA(ctx, st, (len(sub.kf_fixed), len(sub.pt_pos_w), len(sub.line_plucker), len(sub.pt_edge_kf),
                           len(sub.line_edge_kf)), comm=comm)
    tries_per_step = 15  # one full local BA = 5 + 10 LM iterations
    for _ in range(max(warmup, 3)):
        ba.bench_tries(tries_per_step)
    ctx.sync()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    done = 0
    for _ in range(steps):
        it, tr = ba.bench_tries(tries_per_step)
        done += tr
    e1.record(stream)
    ctx.sync()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=torch.cuda.current_device())
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # full solve latency (upload excluded, result download included)
    t0 = time.perf_counter()
    out = ba.solve()
    solve_ms = 1e3 * (time.perf_counter() - t0)
    ba.close()
    if comm:
        comm.close()
    edges = len(prob.pt_edge_kf) + len(prob.line_edge_kf) + len(prob.plane_edge_lm)
    alg_bytes = edges * 32 + len(prob.pt_pos_w) * 48 + len(prob.line_plucker) * 96 + len(prob.kf_fixed) * 112
    return {"metric": "local_ba_lm_iterations_per_sec", "value": done / (ms * 1e-3), "unit": "LM iterations/s",
            "scaling": "strong", "lm_tries_timed": done, "ms_per_lm_iteration": ms / max(done, 1),
            "full_solve_ms": solve_ms, "solve_iters": [out["iters_first"], out["iters_second"], out["lm_tries"]],
            "config": {"workload": "local_bundle_adjuster 20 local + 10 fixed KF, 4000 points + 800 lines + 200 plane edges",
                       "edges": edges, "parallelism": f"landmark-sharded over {world} GPU(s), 1 packed all-reduce per LM try"},
            "algorithmic_bytes_per_iteration": alg_bytes,
            "hbm_roofline_frac": (alg_bytes / (ms * 1e-3 / max(done, 1))) / 1e9 / _peaks()[0]}


def bench_mapping(pkg, ctx, cpu_baseline, reps=10):
    """SURVEY 8(f) rows measured through the C ABI with HOST buffers (the calls the mapping / relocalisation threads make;
    every call includes its H2D/D2H copies): match::fuse search (20 target keyframes x 1000 landmarks,
    mapping_module.cc:711-714), DBoW2 transform on a full-size synthetic vocabulary (k = 10, L = 6, 1 111 110 nodes like the
    shipped orb_vocab.dbow2) and match::bow_tree (one frame x 20 candidate keyframes, relocalizer.cc:79).  Wall clock
    around synchronous calls, median of `reps`; the CPU oracle port (one thread) of the same call beside it."""
    import bow_data
    import fuse_data
    import oracle_api
    import synth
    out = {}

    def med(fn):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return 1e3 * float(np.median(ts))

    orc = oracle_api.Oracle() if cpu_baseline else None
    # ---- fuse
    grid = pkg.capi.make_grid(COLS, ROWS)
    cam = pkg.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, COLS, ROWS)
    sf, isg = synth.scale_factors(), fuse_data.inv_level_sigma_sq()
    lms, targets = fuse_data.make_point_fuse_scene(77, m=1000, num_targets=20)
    l0 = ctx.launch_count()
    ms = med(lambda: ctx.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, targets, lms, 3.0, 1))
    pairs = 20 * 1000
    out["fuse_replace_duplication"] = {"ms_per_call": ms, "pairs_per_s": pairs / (ms * 1e-3), "config": "20 target keyframes x 1000 landmarks, ~1200 keypoints each",
                                       "gpu_launches": int(ctx.launch_count() - l0)}
    if orc:
        t0 = time.perf_counter()
        for t in targets:
            orc.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, t, lms, 3.0, 1)
        out["fuse_replace_duplication"]["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)
    # ---- DBoW2 transform: full-size complete tree (breadth-first ids: parent(i) = (i - 1) // k)
    k, L = 10, 6
    n_nodes = (k ** (L + 1) - 1) // (k - 1)
    rng = np.random.default_rng(5)
    parent = ((np.arange(1, n_nodes, dtype=np.int64) - 1) // k).astype(np.int32)
    vdesc = rng.integers(0, 256, size=(n_nodes - 1, 32), dtype=np.uint8)
    leaf = (np.arange(1, n_nodes) >= (k ** L - 1) // (k - 1)).astype(np.uint8)
    weight = (rng.uniform(0.1, 9.7, n_nodes - 1) * leaf).astype(np.float32)
    vocab = pkg.BowVocabulary(ctx, k=k, L=L, parent=parent, desc=vdesc, weight=weight, is_leaf=leaf)
    rows = 64 * 1000
    desc = rng.integers(0, 256, size=(rows, 32), dtype=np.uint8)
    l0 = ctx.launch_count()
    ms = med(lambda: vocab.transform(desc, 4))
    alg = rows * (32 + 12 + L * k * 32)   # descriptor in, 3 outputs, k child descriptors per level (L2-resident tree)
    out["bow_transform"] = {"ms_per_call": ms, "descriptors_per_s": rows / (ms * 1e-3), "rows": rows,
                            "config": f"k={k} L={L} {n_nodes} nodes ({(n_nodes * 32) >> 20} MB of descriptors), levelsup 4",
                            "algorithmic_bytes": alg, "gathered_GBps": alg / (ms * 1e-3) / 1e9,
                            "gpu_launches": int(ctx.launch_count() - l0)}
    if orc:
        ov = orc.bow_vocab_create(k, L, parent, vdesc, weight, leaf)
        t0 = time.perf_counter()
        orc.bow_transform(ov, desc[:8000], 4)
        out["bow_transform"]["cpu_port_descriptors_per_s"] = 8000 / (time.perf_counter() - t0)
        orc.bow_vocab_destroy(ov)
    vocab.close()
    # ---- match::bow_tree: one frame x 20 candidate keyframes
    frame, bp = None, []
    for sd in range(20):
        s1, s2, _ = bow_data.make_bow_sides(100 + sd, n1=1000, n2=1000, num_nodes=90)
        if frame is None:
            frame = dict(s2)
            frame.pop("valid")
        bp.append((s1, frame))
    l0 = ctx.launch_count()
    ms = med(lambda: ctx.match_bow_tree(bp, 0.75, True))
    out["bow_tree_match"] = {"ms_per_call": ms, "pairs_per_s": 20 / (ms * 1e-3), "config": "1 frame x 20 keyframes, 1000 keypoints, ~110 nodes",
                             "gpu_launches": int(ctx.launch_count() - l0)}
    if orc:
        t0 = time.perf_counter()
        for a, b in bp:
            orc.bow_tree_match(a, b, 0.75, True)
        out["bow_tree_match"]["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)
    # ---- solve::essential_solver RANSAC (robust.cc:232: 50 hypotheses over ~500 brute-force matches)
    try:
        import ess_data
        b1, b2, matches, _ = ess_data.make_two_view(10, n=500, outlier_frac=0.3)
        smp = ess_data.draw_samples(1, len(matches), 50)
        l0 = ctx.launch_count()
        ms = med(lambda: ctx.essential_ransac(b1, b2, matches, smp, False))
        out["essential_ransac"] = {"ms_per_call": ms, "config": "500 matches, 50 eight-point hypotheses",
                                   "gpu_launches": int(ctx.launch_count() - l0)}
        if orc:
            t0 = time.perf_counter()
            orc.essential_ransac(b1, b2, matches, smp, False)
            out["essential_ransac"]["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)
    except Exception as e:
        out["essential_ransac"] = {"error": f"{type(e).__name__}: {e}"}
    # ---- Planar_Mapping_module plane RANSAC (first GPU execution of this path happens at round end)
    try:
        import plane_data
        pts, valid, _, _ = plane_data.make_plane_cloud(10, n=400)
        psmp = plane_data.draw_plane_samples(0, valid, 50, 18)
        l0 = ctx.launch_count()
        ms = med(lambda: ctx.plane_ransac(pts, valid, psmp, plane_data.CFG_ESTIMATE))
        out["plane_ransac"] = {"ms_per_call": ms, "config": "400 landmarks, 50 hypotheses of 18 points",
                               "gpu_launches": int(ctx.launch_count() - l0)}
        if orc:
            want = orc.plane_ransac(pts, valid, psmp, plane_data.CFG_ESTIMATE)
            got = ctx.plane_ransac(pts, valid, psmp, plane_data.CFG_ESTIMATE)
            out["plane_ransac"]["matches_oracle"] = bool(got[0] == want[0] and np.array_equal(got[1], want[1]) and
                                                         np.array_equal(got[3], want[3]))
            t0 = time.perf_counter()
            orc.plane_ransac(pts, valid, psmp, plane_data.CFG_ESTIMATE)
            out["plane_ransac"]["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)
    except Exception as e:
        out["plane_ransac"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def bench_ba_cpu(n_solves=3):
    """CPU oracle port of the same local BA (single thread)."""
    import ba_data
    import oracle_api
    orc = oracle_api.Oracle()
    prob = ba_data.make_ba_problem(42)
    t0 = time.perf_counter()
    tries = 0
    for _ in range(n_solves):
        r = ba_data.oracle_local_ba(orc, prob)
        tries += r.lm_tries
    dt = time.perf_counter() - t0
    return {"value": tries / dt, "unit": "LM iterations/s", "cores": 1, "kind": "port",
            "sample": f"{n_solves} full local-BA solves ({tries} LM tries, {dt:.1f} s)"}


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path (oracle port; the reference binary cannot be built here)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample per step: 2 frames per host thread (capped), same workload/config as the GPU arm
    per_step = int(min(256, max(16, 2 * cores)))
    seqs, frames, t_idx = build_inputs(per_step, args.seed)
    import oracle_api
    import synth
    orc = oracle_api.Oracle()
    p = oracle_api.orb_params()
    rng = np.random.default_rng(args.seed)
    with ThreadPoolExecutor(max_workers=cores) as ex:
        kps = list(ex.map(lambda st: orc.orb_extract(p, seqs[st[0]].frames[st[1] - 1]), t_idx))
    lasts = [seqs[s].last_frame_landmarks(t - 1, kps[b]["kps"], kps[b]["desc"]) for b, (s, t) in enumerate(t_idx)]
    preds = np.stack([seqs[s].predicted_pose(t, rng) for (s, t) in t_idx])
    aux = dict(seqs=seqs, t_idx=t_idx, lasts=lasts, preds=preds)
    for _ in range(max(args.warmup, 1)):
        cpu_port_frames(frames, aux, list(range(min(per_step, cores))), cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_frames(frames, aux, list(range(per_step)), cores)
    dt = time.perf_counter() - t0
    fps = args.steps * per_step / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": per_step,
                   "note": "reference binary unbuildable here (no C++ OpenCV/Eigen/g2o); CPU oracle port, frames spread over host threads"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} frames/step x {args.steps} steps"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=512,
                    help="frames per step per GPU (512 x 307 KB > 126 MB L2); measured r01f: 2 x 296 (whole waves of the "
                         "one-CTA-per-frame kernels) is 1 % slower than 2 x 256 -- the tracking streams already overlap")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--streams", type=int, default=2, help="sub-batches in flight per GPU (one context/stream each)")
    ap.add_argument("--track-streams", type=int, default=1,
                    help="1: matcher / pose optimiser of every sub-batch on a high-priority stream of its own")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ba", action="store_true", help="skip the local-BA metric")
    ap.add_argument("--no-lines", action="store_true", help="skip the LSD+LBD line front-end metric")
    ap.add_argument("--only-lines", action="store_true", help="development: run only the line front-end leg")
    ap.add_argument("--no-mapping", action="store_true", help="skip the fuse / BoW legs (SURVEY 8(f) rows)")
    ap.add_argument("--no-stereo", action="store_true", help="skip the stereo point+line front-end leg (configs[4])")
    ap.add_argument("--only-stereo", action="store_true", help="development: run only the stereo leg")
    ap.add_argument("--stereo-batch", type=int, default=148, help="stereo frames per step per GPU")
    ap.add_argument("--line-batch", type=int, default=1776, help="frames per step per GPU of the line front-end leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    pkg = _load_pkg()
    ctx = pkg.Context(local_rank)
    lib = pkg.lib()
    B = args.batch
    stream = torch.cuda.ExternalStream(lib.plp_ctx_stream(ctx.handle), device=f"cuda:{local_rank}")
    if args.only_lines:
        r = bench_lines(pkg, ctx, stream, rank, world, args.steps, args.warmup, args.line_batch, args.seed,
                        world == 1 and not args.no_cpu_baseline)
        if rank == 0:
            print(json.dumps(r))
        return
    if args.only_stereo:
        r = bench_stereo(pkg, ctx, stream, rank, world, args.steps, args.warmup, args.stereo_batch, args.seed)
        if rank == 0:
            print(json.dumps(r))
        return
    # The batch of a step is split into `--streams` sub-batches, each owned by its own context (= CUDA stream) with its
    # own extractor / tracker handles: the H2D copy, the one-CTA-per-frame matcher / pose optimiser and the D2H copy of
    # one sub-batch overlap the extraction kernels of the other (plain stream concurrency, no graph capture).
    S = max(1, min(args.streams, B))
    Bs = B // S
    B = Bs * S
    ctxs = [ctx] + [pkg.Context(local_rank) for _ in range(S - 1)]
    # tracking (one CTA per frame: matcher, pose optimiser) on high-priority streams of its own
    tctxs = [pkg.Context(local_rank, high_priority=True) for _ in range(S)] if args.track_streams else [None] * S
    fes, frames_l, auxs = [], [], []
    for c in range(S):
        fe_c, fr_c, aux_c = setup_front_end(pkg, ctxs[c], Bs, args.seed + 1000 * rank + 37 * c, tctxs[c])
        fes.append(fe_c)
        frames_l.append(fr_c)
        auxs.append(aux_c)
    fe, frames, aux = fes[0], np.concatenate(frames_l), auxs[0]

    all_ctxs = ctxs + [t for t in tctxs if t is not None]

    def barrier():
        for cx in all_ctxs:
            cx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def join_streams():  # stream 0 waits for the work enqueued on the other streams
        for cx in all_ctxs[1:]:
            ctx._check(lib.plp_ctx_wait_ctx(ctx.handle, cx.handle))

    def step_all():
        for c in range(S):
            fes[c].step(Bs)

    # ---------------- value: device-resident ------------------------------------------------------------
    for _ in range(args.warmup):
        step_all()
    barrier()
    launches0 = sum(cx.launch_count() for cx in all_ctxs)
    with ClockSampler(local_rank) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_all()
        join_streams()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
    launches = sum(cx.launch_count() for cx in all_ctxs) - launches0
    ok = 0
    for c in range(S):
        ok += int((fes[c].download_tracking(Bs)["num_valid"] >= 20).sum())
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---------------- e2e: host buffers, copies inside the timed region ----------------------------------
    pinned_in, pinned_out = [], []
    out_bytes = Bs * (128 + 4 + 4)
    for c in range(S):
        pi, po = C.c_void_p(), C.c_void_p()
        ctx._check(lib.plp_host_alloc_pinned(C.c_size_t(frames_l[c].nbytes), C.byref(pi)))
        ctx._check(lib.plp_host_alloc_pinned(C.c_size_t(out_bytes), C.byref(po)))
        C.memmove(pi, frames_l[c].ctypes.data, frames_l[c].nbytes)
        pinned_in.append(pi)
        pinned_out.append(po)
    d2h_bytes = S * out_bytes

    def e2e_step():
        for c in range(S):
            cx, f = ctxs[c], fes[c]
            cx._check(lib.plp_dev_upload_async(cx.handle, f.d_imgs.ptr, pinned_in[c], C.c_size_t(frames_l[c].nbytes)))
            f.step(Bs)
            po = pinned_out[c].value
            cx = f.track_ctx  # the results are produced on the tracking stream
            cx._check(lib.plp_dev_download_async(cx.handle, C.c_void_p(po), f.d_pose.ptr, C.c_size_t(Bs * 128)))
            cx._check(lib.plp_dev_download_async(cx.handle, C.c_void_p(po + Bs * 128), f.d_num_valid.ptr, C.c_size_t(Bs * 4)))
            cx._check(lib.plp_dev_download_async(cx.handle, C.c_void_p(po + Bs * 132), f.d_n_inl.ptr, C.c_size_t(Bs * 4)))

    for _ in range(2):
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(args.steps):
        e2e_step()
    join_streams()
    f1.record(stream)
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    # the results of the last step are in host memory now: read them (the "loss or metric" of the contract)
    e2e_ok = 0
    for c in range(S):
        nv = np.frombuffer((C.c_char * (Bs * 4)).from_address(pinned_out[c].value + Bs * 128), np.int32)
        e2e_ok += int((nv >= 20).sum())
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t.item()) * 1e-3)

    # ---------------- roofline: per-kernel event timing over the same steps ------------------------------
    tim_ctxs = [fe.ctx] + ([fe.track_ctx] if fe.track_ctx is not fe.ctx else [])
    for cx in tim_ctxs:
        cx._check(lib.plp_ctx_kernel_timing(cx.handle, 1))
    for _ in range(args.steps):
        fe.step(Bs)
    kt = {}
    for cx in tim_ctxs:
        buf = C.create_string_buffer(1 << 16)
        cx._check(lib.plp_ctx_kernel_timing_report(cx.handle, buf, C.c_size_t(len(buf))))
        cx._check(lib.plp_ctx_kernel_timing(cx.handle, 0))
        kt.update(json.loads(buf.value.decode()))
    total_ms = sum(v["total_ms"] for v in kt.values())
    shares = {k: round(v["total_ms"] / total_ms, 4) for k, v in sorted(kt.items(), key=lambda kv: -kv[1]["total_ms"])}
    dom = max(kt.items(), key=lambda kv: kv[1]["total_ms"])
    peak, peak_src = _peaks()
    dom_name = dom[0]
    per_launch_ms = dom[1]["total_ms"] / dom[1]["count"]
    launches_per_step = dom[1]["count"] / max(args.steps, 1)
    alg = ALG_BYTES.get(dom_name)
    if alg is None:
        alg = PYR_PX
    # algorithmic bytes of ONE step of this kernel (all of its launches in a step together process the sub-batch once:
    # the matcher's second launch only revisits the few frames that need the wider margin) over its time per step
    alg_bytes_launch = alg * Bs
    per_step_ms = dom[1]["total_ms"] / max(args.steps, 1)
    achieved = alg_bytes_launch / (per_step_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tf = sorted((ROOT / "profiles").glob("traffic_*.json"))
    if tf:  # dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full` capture, scaled to this launch
        tj = json.loads(tf[-1].read_text())
        kk = tj["kernels"].get(dom_name) or tj["kernels"].get(dom_name.replace("_v2", ""))
        if kk:
            traffic = (kk["dram_bytes_read"] + kk["dram_bytes_write"]) / tj["frames_per_launch"] * Bs
            traffic_src = f"{tf[-1].name}: {tj['source']}; per-frame bytes x {Bs} frames"
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_launch, "ms_per_launch": per_launch_ms,
                "launches_per_step": launches_per_step, "ms_per_step_of_this_kernel": per_step_ms,
                "kernel_time_shares": shares,
                "frames_per_launch": Bs,
                "how": "CUDA events around every launch on the launching stream over a repeat of the timed steps "
                       "(one sub-batch, kernels serialised)"}

    ba_res = None
    if not args.no_ba:
        ba_res = bench_ba(pkg, ctx, stream, rank, world, steps=max(args.steps, 5), warmup=args.warmup)

    lines_res = None
    if not args.no_lines:
        lines_res = bench_lines(pkg, ctx, stream, rank, world, args.steps, args.warmup, args.line_batch, args.seed,
                                world == 1 and not args.no_cpu_baseline)

    stereo_res = None
    if not args.no_stereo:
        stereo_res = bench_stereo(pkg, ctx, stream, rank, world, max(3, args.steps // 2), args.warmup, args.stereo_batch,
                                  args.seed)

    mapping_res = None
    if rank == 0 and world == 1 and not args.no_mapping:
        try:
            mapping_res = bench_mapping(pkg, ctx, not args.no_cpu_baseline)
        except Exception as e:  # an auxiliary leg must not take the headline line down with it; say so loudly
            mapping_res = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B, "image": f"{COLS}x{ROWS}",
                       "orb": {"max_num_keypts": 1000, "scale_factor": 1.2, "num_levels": 8, "ini_fast_thr": 20, "min_fast_thr": 7},
                       "parallelism": f"frames sharded over {world} GPU(s), no data-path collective",
                       "streams_per_gpu": S * (2 if args.track_streams else 1), "frames_per_stream_per_step": Bs,
                       "stream_layout": ("per sub-batch: extraction stream + high-priority tracking stream"
                                         if args.track_streams else "per sub-batch: one stream"),
                       "l2": "inputs larger than L2 (batch x 307 KB images)",
                       "tracked_ok_frames": ok, "tracked_ok_frames_e2e": e2e_ok},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(frames.nbytes), "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_sample = int(min(Bs, max(32, 2 * cores)))
            fps_mt, dt_mt, _ = cpu_port_frames(frames_l[0], aux, list(range(n_sample)), cores)
            fps_1, dt_1, _ = cpu_port_frames(frames_l[0], aux, list(range(min(8, n_sample))), 1)
            line["cpu_baseline"] = {"value": fps_mt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_sample} frames of the same workload over {cores} host threads ({dt_mt:.1f} s)",
                                    "single_thread_value": fps_1}
        if ba_res is not None:
            if world == 1 and not args.no_cpu_baseline:
                ba_res["cpu_baseline"] = bench_ba_cpu()
            line["local_ba"] = ba_res
        if lines_res is not None:
            line["line_frontend"] = lines_res
        if stereo_res is not None:
            line["stereo_frontend"] = stereo_res
        if mapping_res is not None:
            line["mapping_matchers"] = mapping_res
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
