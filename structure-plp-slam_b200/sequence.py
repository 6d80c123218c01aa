"""One live sequence through the host entry points, in the call order of tracking_module::track.

This is the "config 1" plumbing loop of SURVEY.md section 8(d) (example/run_tum_rgbd_slam.cc:84-122 feeding
tracking_module.cc:424-570): per frame

    orb_extractor::extract                                   (data/frame.cc:1125-1140)
    frame_tracker::motion_based_track                        (module/frame_tracker.cc:52-124)
        projection::match_current_and_last_frames(margin 20, retried with 40 below 20 matches)
        pose_optimizer::optimize, discard_outliers
    tracking_module::optimize_current_frame_with_local_map   (tracking_module.cc:732-759)
        search_local_landmarks -> projection::match_frame_and_landmarks(margin 5, ratio 0.8)   (:908-984)
        pose_optimizer::optimize

with frame t depending on the pose estimated for frame t-1 (motion model, frame_tracker.cc:58-59).  The host-side
bookkeeping the reference keeps in data::frame / data::landmark (which keypoint carries which landmark, the local
map, keyframe insertion) is modelled here with plain arrays; every COMPUTE step is one call into a backend that has the
reference's method names.  `GpuBackend` routes them to the C ABI (host buffers in, host buffers out, each call
synchronous like the reference's); the tests run the same loop with the CPU oracle behind the same interface and
compare trajectories.  No compute happens in this file.
"""
from __future__ import annotations

import time

import numpy as np

from .capi import PT_OBS_DTYPE, Context, OrbExtractor, make_grid

NUM_MATCHES_THR = 20      # frame_tracker::num_matches_thr_
LOCAL_MAP_MARGIN = 5.0    # tracking_module.cc:976-981 (monocular, not recently relocalised)
LOCAL_MAP_RATIO = 0.8     # tracking_module.cc:975 match::projection projection_matcher(0.8)


class GpuBackend:
    """The reference's operator calls on the B200 through libplpslam_b200.so."""

    def __init__(self, pkg, ctx: Context, rows: int, cols: int, max_num_keypts=1000):
        self.ctx = ctx
        self.orb = OrbExtractor(ctx, rows, cols, max_num_keypts, max_batch=1)
        self.scale_factors = np.asarray(self.orb.scale_factors, np.float32)
        self.inv_level_sigma_sq = np.asarray(self.orb.inv_level_sigma_sq, np.float32)

    def extract(self, img):
        kps, desc = self.orb.extract(img)
        return kps, desc

    def match_current_and_last_frames(self, grid, cam, curr, Tc, Tl, last, margin):
        return self.ctx.match_current_and_last_frames(grid, self.scale_factors, cam, curr, Tc, Tl, last, margin, True)

    def match_frame_and_landmarks(self, grid, frm, q, margin, lowe_ratio):
        return self.ctx.match_frame_and_landmarks(grid, self.scale_factors, frm, q, margin, lowe_ratio)

    def pose_optimize(self, cam, T, pts):
        T_out, pt_out, _, n_inl = self.ctx.pose_optimize(cam, T, pts)
        return T_out, pt_out, n_inl

    def close(self):
        self.orb.close()


def _project(cam, T, X):
    """camera::perspective::reproject_to_image (camera/perspective.cc:190-209) for the host-side can_observe model."""
    Xc = X @ T[:3, :3].T + T[:3, 3]
    z = Xc[:, 2]
    ok = z > 0
    zi = 1.0 / np.where(ok, z, 1.0)
    u = cam.fx * Xc[:, 0] * zi + cam.cx
    v = cam.fy * Xc[:, 1] * zi + cam.cy
    ok &= (cam.min_x < u) & (u < cam.max_x) & (cam.min_y < v) & (v < cam.max_y)
    return u, v, ok


class SequentialTracker:
    def __init__(self, backend, rows, cols, cam, keyframe_every=4, local_keyframes=3):
        self.be = backend
        self.cam = cam
        self.grid = make_grid(cols, rows)
        self.keyframe_every = keyframe_every
        self.local_keyframes = local_keyframes

    @classmethod
    def for_gpu(cls, pkg, ctx, rows, cols, cam, with_lines=False, **kw):
        return cls(GpuBackend(pkg, ctx, rows, cols), rows, cols, cam, **kw)

    def close(self):
        if hasattr(self.be, "close"):
            self.be.close()

    # ------------------------------------------------------------------------------------------------------
    def _pts(self, kps, idx, pos_w):
        pts = np.zeros(len(idx), PT_OBS_DTYPE)
        pts["pos_w"] = pos_w
        pts["obs_x"], pts["obs_y"] = kps["x"][idx], kps["y"][idx]
        pts["x_right"] = -1.0
        pts["inv_sigma_sq"] = self.be.inv_level_sigma_sq[kps["octave"][idx]]
        return pts

    def run(self, seq, trace=None):
        """seq: an object with .frames[t] (uint8 images), .poses[0] (initial pose) and .backproject(T, x, y) (the map
        initialisation / triangulation stand-in: keypoints become landmarks on the known scene geometry)."""
        be, cam, grid = self.be, self.cam, self.grid
        n_frames = len(seq.frames)
        # ---- map: landmark arrays (position, descriptor, creation octave, creation keyframe)
        lm_pos = np.zeros((0, 3))
        lm_desc = np.zeros((0, 32), np.uint8)
        lm_oct = np.zeros(0, np.int32)
        lm_kf = np.zeros(0, np.int32)
        poses = [np.array(seq.poses[0], np.float64)]
        stage_ms, tracked = [], 0
        n_kf = 0

        def add_keyframe(kps, desc, lm_of_kp, T):
            nonlocal lm_pos, lm_desc, lm_oct, lm_kf, n_kf
            new = np.nonzero(lm_of_kp < 0)[0]
            X = seq.backproject(T, kps["x"][new].astype(np.float64), kps["y"][new].astype(np.float64))
            lm_of_kp[new] = len(lm_pos) + np.arange(len(new))
            lm_pos = np.concatenate([lm_pos, X])
            lm_desc = np.concatenate([lm_desc, desc[new]])
            lm_oct = np.concatenate([lm_oct, kps["octave"][new].astype(np.int32)])
            lm_kf = np.concatenate([lm_kf, np.full(len(new), n_kf, np.int32)])
            n_kf += 1

        # frame 0 = the initial keyframe
        t0 = time.perf_counter()
        kps, desc = be.extract(seq.frames[0])
        ms = {"extract": 1e3 * (time.perf_counter() - t0), "motion_match": 0.0, "pose_opt_1": 0.0, "local_map_match": 0.0,
              "pose_opt_2": 0.0}
        stage_ms.append(ms)
        lm_of_kp = np.full(len(kps), -1, np.int64)
        add_keyframe(kps, desc, lm_of_kp, poses[0])
        last = dict(kps=kps, desc=desc, lm=lm_of_kp, pose=poses[0])
        velocity = np.eye(4)
        for t in range(1, n_frames):
            ms = {}
            c0 = time.perf_counter()
            kps, desc = be.extract(seq.frames[t])
            c1 = time.perf_counter()
            ms["extract"] = 1e3 * (c1 - c0)
            curr = dict(x=kps["x"], y=kps["y"], octave=kps["octave"], angle=kps["angle"], desc=desc)
            # ---- motion-based track (frame_tracker.cc:52-124)
            T_pred = velocity @ last["pose"]
            has = np.nonzero(last["lm"] >= 0)[0]
            lastd = dict(pos_w=lm_pos[last["lm"][has]], octave=last["kps"]["octave"][has].astype(np.int32),
                         angle=last["kps"]["angle"][has].astype(np.float32), desc=last["desc"][has],
                         valid=np.ones(len(has), np.uint8))
            matched, nm = be.match_current_and_last_frames(grid, cam, curr, T_pred, last["pose"], lastd, 20.0)
            if nm < NUM_MATCHES_THR:
                matched, nm = be.match_current_and_last_frames(grid, cam, curr, T_pred, last["pose"], lastd, 40.0)
            c2 = time.perf_counter()
            ms["motion_match"] = 1e3 * (c2 - c1)
            lm_of_kp = np.full(len(kps), -1, np.int64)
            T_cur = T_pred
            if nm >= NUM_MATCHES_THR:
                idx = np.nonzero(matched >= 0)[0]
                lm_of_kp[idx] = last["lm"][has[matched[idx]]]
                T_cur, pt_out, n_inl = be.pose_optimize(cam, T_pred, self._pts(kps, idx, lm_pos[lm_of_kp[idx]]))
                lm_of_kp[idx[np.asarray(pt_out) != 0]] = -1   # discard_outliers (frame_tracker.cc:253-283)
            c3 = time.perf_counter()
            ms["pose_opt_1"] = 1e3 * (c3 - c2)
            # ---- local map (tracking_module.cc:908-984): landmarks of the last keyframes not yet seen in this frame
            local = np.nonzero(lm_kf >= n_kf - self.local_keyframes)[0]
            seen = np.zeros(len(lm_pos), bool)
            seen[lm_of_kp[lm_of_kp >= 0]] = True
            local = local[~seen[local]]
            u, v, ok = _project(cam, T_cur, lm_pos[local])
            q = dict(reproj_x=u.astype(np.float32), reproj_y=v.astype(np.float32), x_right=np.full(len(local), -1.0, np.float32),
                     scale_level=lm_oct[local], desc=lm_desc[local], valid=ok.astype(np.uint8))
            frm = dict(curr, claimed=(lm_of_kp >= 0).astype(np.uint8))
            best, nb = be.match_frame_and_landmarks(grid, frm, q, LOCAL_MAP_MARGIN, LOCAL_MAP_RATIO)
            c4 = time.perf_counter()
            ms["local_map_match"] = 1e3 * (c4 - c3)
            hit = np.nonzero(best >= 0)[0]
            lm_of_kp[best[hit]] = local[hit]
            idx = np.nonzero(lm_of_kp >= 0)[0]
            n_inl = 0
            if len(idx) >= 5:
                T_cur, pt_out, n_inl = be.pose_optimize(cam, T_cur, self._pts(kps, idx, lm_pos[lm_of_kp[idx]]))
                lm_of_kp[idx[np.asarray(pt_out) != 0]] = -1
            c5 = time.perf_counter()
            ms["pose_opt_2"] = 1e3 * (c5 - c4)
            stage_ms.append(ms)
            if trace is not None:
                trace.append(dict(t=t, n_kp=len(kps), n_last=len(has), motion_matches=int(nm), local_queries=len(local),
                                  local_matches=int(nb), inliers=int(n_inl), matched=matched.copy(), best=best.copy()))
            tracked += int(n_inl >= NUM_MATCHES_THR)
            # motion model update (tracking_module.cc update_motion_model): velocity = T_cur * T_last^-1
            velocity = T_cur @ np.linalg.inv(last["pose"])
            poses.append(T_cur)
            if t % self.keyframe_every == 0:   # new keyframe: untracked keypoints become landmarks (mapping stand-in)
                add_keyframe(kps, desc, lm_of_kp, T_cur)
            last = dict(kps=kps, desc=desc, lm=lm_of_kp, pose=T_cur)
        return dict(poses=poses, stage_ms=stage_ms, tracked=tracked)
