"""Seeded scenes for the match::fuse parity tests, and a small map-state model of the reference's sequential effects
(landmark::add_observation / replace / compute_descriptor) used to check the adapter protocol of INTEGRATION.md:
"search the whole (target keyframe x landmark) batch at once, apply the effects in the reference's order, re-issue the
search for landmarks whose descriptor was recomputed" must equal the reference's landmark-by-landmark loop."""
from __future__ import annotations

import numpy as np

import synth

LOG_SF = float(np.log(np.float32(1.2)).astype(np.float32))  # frame::log_scale_factor_ = std::log(scale_factor_) in float


def inv_level_sigma_sq(num_levels=8, sf=1.2):
    s = synth.scale_factors(num_levels, sf)
    return (np.float32(1.0) / (s * s)).astype(np.float32)


def _pose_parts(T):
    R = np.ascontiguousarray(T[:3, :3], np.float64)
    t = np.ascontiguousarray(T[:3, 3], np.float64)
    return R, t, -R.T @ t


def make_point_fuse_scene(seed, m=800, num_targets=4, n_extra=250, stereo=False, num_levels=8, obs_frac=0.8):
    """Landmarks + `num_targets` target keyframes that re-observe most of them (noisy, on the level grid)."""
    rng = np.random.default_rng(seed)
    sf = synth.scale_factors(num_levels)
    X = np.stack([rng.uniform(-4, 4, m), rng.uniform(-3, 3, m), rng.uniform(2, 12, m)], 1)
    desc = synth.rand_desc(rng, m)
    ref_center = rng.normal(0, 0.2, 3)
    v = X - ref_center
    dist0 = np.linalg.norm(v, axis=1)
    normal = v / dist0[:, None] + rng.normal(0, 0.15, (m, 3))
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    flip = rng.random(m) < 0.08  # fail the 60-degree gate
    normal[flip] *= -1
    min_raw = (dist0 * rng.uniform(0.3, 1.2, m)).astype(np.float32)       # some fail the min gate
    max_raw = (np.maximum(dist0, min_raw) * rng.uniform(0.7, 3.4, m)).astype(np.float32)  # some fail the max gate
    lms = dict(pos_w=X, obs_mean_normal=normal, desc=desc, max_valid_dist_raw=max_raw,
               min_valid_dist=(0.7 * min_raw.astype(np.float64)).astype(np.float32),   # landmark.cc:297-301
               max_valid_dist=(1.3 * max_raw.astype(np.float64)).astype(np.float32),   # landmark.cc:303-307
               valid=(rng.random(m) > 0.07).astype(np.uint8))
    targets = []
    for _ in range(num_targets):
        T = synth.make_pose(rng, 0.03, 0.15)
        R, t, c = _pose_parts(T)
        uv, z = synth.project(T, X)
        d = np.linalg.norm(X - c, axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            lvl = np.ceil(np.log(max_raw / d.astype(np.float32)) / np.float32(LOG_SF))
        lvl = np.clip(np.nan_to_num(lvl, nan=0.0), 0, num_levels - 1).astype(np.int32)
        octv = np.clip(lvl - rng.integers(0, 3, m) + rng.integers(0, 2, m), 0, num_levels - 1).astype(np.int32)
        noise = rng.normal(0, 0.9, (m, 2)) * sf[octv][:, None]
        big = rng.random(m) < 0.1     # fail the chi-square gate
        noise[big] *= 6
        s = sf[octv][:, None].astype(np.float32)
        pts = (np.round((uv + noise) / s) * s).astype(np.float32)
        kdesc = synth.flip_bits(rng, desc, rng.integers(3, 75, m))
        keep = (rng.random(m) > 1.0 - obs_frac) & (z > 0.1)
        xs, ys, octs, descs = [pts[keep, 0]], [pts[keep, 1]], [octv[keep]], [kdesc[keep]]
        zs = [z[keep]]
        nd = int(0.15 * keep.sum())
        src = rng.choice(np.nonzero(keep)[0], nd, replace=False)
        dp = pts[src] + (rng.integers(-1, 2, (nd, 2)) * sf[octv[src]][:, None]).astype(np.float32)
        dd = kdesc[src].copy()
        mod = rng.random(nd) < 0.5     # the other half are exact descriptor ties
        dd[mod] = synth.flip_bits(rng, dd[mod], rng.integers(1, 4, mod.sum()))
        xs.append(dp[:, 0]); ys.append(dp[:, 1]); octs.append(octv[src]); descs.append(dd); zs.append(z[src])
        xs.append(rng.uniform(-5, synth.COLS + 5, n_extra).astype(np.float32))
        ys.append(rng.uniform(-5, synth.ROWS + 5, n_extra).astype(np.float32))
        octs.append(rng.integers(0, num_levels, n_extra).astype(np.int32))
        descs.append(synth.rand_desc(rng, n_extra)); zs.append(rng.uniform(2, 12, n_extra))
        x = np.concatenate(xs).astype(np.float32)
        perm = rng.permutation(len(x))
        tgt = dict(x=x[perm], y=np.concatenate(ys).astype(np.float32)[perm],
                   octave=np.concatenate(octs).astype(np.int32)[perm], desc=np.concatenate(descs)[perm],
                   rot_cw=R, trans_cw=t, cam_center=c, skip=(rng.random(m) < 0.1).astype(np.uint8))
        if stereo:
            zz = np.concatenate(zs)[perm]
            xr = (tgt["x"] - synth.BF / zz + rng.normal(0, 0.7, len(x))).astype(np.float32)
            xr[rng.random(len(x)) < 0.3] = -1.0
            tgt["x_right"] = xr
        targets.append(tgt)
    return lms, targets


def make_line_fuse_scene(seed, m=250, num_targets=3, n_extra=60, num_levels=1):
    rng = np.random.default_rng(seed)
    P0 = np.stack([rng.uniform(-4, 4, m), rng.uniform(-3, 3, m), rng.uniform(2, 10, m)], 1)
    d = rng.normal(0, 1, (m, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    P1 = P0 + d * rng.uniform(0.5, 3.0, (m, 1))
    far = rng.random(m) < 0.12   # end point behind the camera / out of the image: partial-occlusion branch
    P1[far, 2] -= 14
    desc = synth.rand_desc(rng, m)
    dist0 = np.linalg.norm(0.5 * (P0 + P1), axis=1)
    min_raw = (dist0 * rng.uniform(0.2, 1.1, m)).astype(np.float32)
    max_raw = (np.maximum(dist0, min_raw) * rng.uniform(0.8, 4.5, m)).astype(np.float32)
    lms = dict(pos_w=np.concatenate([P0, P1], 1), desc=desc, max_valid_dist_raw=max_raw,
               min_valid_dist=(0.8 * min_raw.astype(np.float64)).astype(np.float32),   # landmark_line.cc:354-358
               max_valid_dist=(1.2 * max_raw.astype(np.float64)).astype(np.float32),   # landmark_line.cc:360-364
               valid=(rng.random(m) > 0.07).astype(np.uint8))
    targets = []
    for _ in range(num_targets):
        T = synth.make_pose(rng, 0.02, 0.08)
        R, t, c = _pose_parts(T)
        a, za = synth.project(T, P0)
        b, zb = synth.project(T, P1)
        keep = (rng.random(m) > 0.2) & (za > 0.1) & (zb > 0.1)
        k = keep.sum()
        sg = np.where(rng.random(k) < 0.12, 5.0, 0.8)     # some fail the chi-square gate
        sx = a[keep, 0] + rng.normal(0, 1, k) * sg
        sy = a[keep, 1] + rng.normal(0, 1, k) * sg
        ex = b[keep, 0] + rng.normal(0, 1, k) * sg
        ey = b[keep, 1] + rng.normal(0, 1, k) * sg
        kd = synth.flip_bits(rng, desc[keep], rng.integers(3, 75, k))
        nd = k // 5
        src = rng.integers(0, k, nd)
        sx = np.concatenate([sx, sx[src] + rng.normal(0, 0.3, nd), rng.uniform(0, synth.COLS, n_extra)])
        sy = np.concatenate([sy, sy[src] + rng.normal(0, 0.3, nd), rng.uniform(0, synth.ROWS, n_extra)])
        ex = np.concatenate([ex, ex[src] + rng.normal(0, 0.3, nd), rng.uniform(0, synth.COLS, n_extra)])
        ey = np.concatenate([ey, ey[src] + rng.normal(0, 0.3, nd), rng.uniform(0, synth.ROWS, n_extra)])
        kd = np.concatenate([kd, kd[src], synth.rand_desc(rng, n_extra)])
        n = len(sx)
        perm = rng.permutation(n)
        targets.append(dict(sx=sx[perm].astype(np.float32), sy=sy[perm].astype(np.float32),
                            ex=ex[perm].astype(np.float32), ey=ey[perm].astype(np.float32),
                            octave=rng.integers(0, num_levels, n).astype(np.int32), desc=kd[perm],
                            rot_cw=R, trans_cw=t, cam_center=c, skip=(rng.random(m) < 0.1).astype(np.uint8)))
    return lms, targets


# ------------------------------------------------------------------------------------------------------------------
# map-state model of the effects (data/landmark.cc:118-161 add/erase_observation, :181-247 compute_descriptor,
# :392-432 replace; data/keyframe.cc add_landmark / replace_landmark / erase_landmark_with_index)
# ------------------------------------------------------------------------------------------------------------------
class MapModel:
    """Keyframes 0..K-1 with per-keypoint landmark slots; landmarks with observation maps and a median descriptor."""

    def __init__(self, kf_descs, lm_desc, observations):
        self.kf_descs = kf_descs                                   # list of (n_k, 32) uint8
        self.kf_lms = [np.full(len(d), -1, np.int64) for d in kf_descs]
        self.desc = lm_desc.copy()                                 # (M, 32)
        self.obs = [dict() for _ in range(len(lm_desc))]           # kf -> idx (std::map<keyframe*, unsigned>)
        self.erased = np.zeros(len(lm_desc), bool)
        self.desc_version = np.zeros(len(lm_desc), np.int64)
        for lm, kf, idx in observations:
            self.obs[lm][kf] = idx
            self.kf_lms[kf][idx] = lm

    def num_observations(self, lm):
        return len(self.obs[lm])

    def is_observed_in_keyframe(self, lm, kf):
        return kf in self.obs[lm]

    def compute_descriptor(self, lm):
        # landmark.cc:181-247: the observation descriptor with the smallest median Hamming distance to the others
        items = sorted(self.obs[lm].items())
        if not items:
            return
        D = np.stack([self.kf_descs[kf][idx] for kf, idx in items])
        bits = np.unpackbits(D, axis=1).astype(np.int32)
        ham = (bits[:, None, :] != bits[None, :, :]).sum(2)
        med = np.sort(ham, axis=1)[:, int(0.5 * (len(items) - 1))]
        new = D[int(np.argmin(med))]
        if not np.array_equal(new, self.desc[lm]):
            self.desc[lm] = new
            self.desc_version[lm] += 1

    def add_observation(self, lm, kf, idx):
        if kf in self.obs[lm]:
            return
        self.obs[lm][kf] = idx
        self.kf_lms[kf][idx] = lm

    def replace(self, this, lm):
        # landmark::replace(lm): `this` is erased, its observations move to `lm`
        if this == lm:
            return
        observations = dict(self.obs[this])
        self.obs[this].clear()
        self.erased[this] = True
        for kf, idx in sorted(observations.items()):
            if kf not in self.obs[lm]:
                self.kf_lms[kf][idx] = lm
                self.obs[lm][kf] = idx
            else:
                self.kf_lms[kf][idx] = -1
        self.compute_descriptor(lm)

    def apply(self, kf, lm, best_idx):
        """fuse.cc:284-318 for one landmark whose search returned best_idx >= 0."""
        other = int(self.kf_lms[kf][best_idx])
        if other >= 0:
            if not self.erased[other]:
                if self.num_observations(lm) < self.num_observations(other):
                    self.replace(lm, other)
                else:
                    self.replace(other, lm)
        else:
            self.add_observation(lm, kf, best_idx)

    def state(self):
        return ([a.copy() for a in self.kf_lms], [dict(o) for o in self.obs], self.erased.copy(), self.desc.copy())


def fuse_sequential(model, lm_ids, kf_order, search_one):
    """The reference's loops (mapping_module.cc:711-714 x fuse.cc:161-300): per keyframe, per landmark, search with the
    CURRENT state, apply.  search_one(kf, lm_ids, descs) -> best_idx per landmark (state-free search)."""
    num_fused = 0
    for kf in kf_order:
        for j, lm in enumerate(lm_ids):
            if model.erased[lm] or model.is_observed_in_keyframe(lm, kf):
                continue
            best = int(search_one(kf, [j], model.desc[[lm]])[0])
            if best < 0:
                continue
            model.apply(kf, lm, best)
            num_fused += 1
    return num_fused


def fuse_batched(model, lm_ids, kf_order, search_batch):
    """The adapter protocol: ONE batched search up front; effects applied in the reference's order with the gates
    re-checked at apply time; landmarks whose descriptor changed are re-searched against the remaining keyframes.
    search_batch(kfs, js, descs) -> best_idx[len(kfs), len(js)]."""
    lm_ids = list(lm_ids)
    best = np.asarray(search_batch(list(kf_order), list(range(len(lm_ids))), model.desc[lm_ids]))
    seen_version = model.desc_version[lm_ids].copy()
    num_fused = 0
    researches = 0
    for pos, kf in enumerate(kf_order):
        for j, lm in enumerate(lm_ids):
            if model.erased[lm] or model.is_observed_in_keyframe(lm, kf):
                continue
            if model.desc_version[lm] != seen_version[j]:
                # descriptor recomputed by an earlier replace(): refresh this landmark's row for the remaining targets
                rest = list(kf_order[pos:])
                best[pos:, j] = np.asarray(search_batch(rest, [j], model.desc[[lm]]))[:, 0]
                seen_version[j] = model.desc_version[lm]
                researches += 1
            b = int(best[pos, j])
            if b < 0:
                continue
            model.apply(kf, lm, b)
            num_fused += 1
    return num_fused, researches


# ------------------------------------------------------------------------------------------------------------------
# independent pure-Python restatement of the fuse search (pins the C++ oracle with a second implementation)
# ------------------------------------------------------------------------------------------------------------------
def _logf(x):
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    return np.float32(libm.logf(float(x)))


def _cv_floor(v):
    i = int(v)
    return i - (i > v)


def _cv_ceil(v):
    i = int(v)
    return i + (i < v)


def fuse_search_points_python(grid, cam, scale_factors, inv_level_sigma_sq, log_scale_factor, tgt, lms, margin, mode):
    """match/fuse.cc:40-151 (mode 0) / :153-300 (mode 1), written from the reference text with numpy scalars carrying the
    reference's types (double reprojection, float window, unsigned / int level gates, float-double chi-square mix)."""
    f32, f64 = np.float32, np.float64
    x, y, octv = tgt["x"].astype(f32), tgt["y"].astype(f32), tgt["octave"]
    xr_all = tgt.get("x_right")
    n, m = len(x), len(lms["min_valid_dist"])
    # assign_keypoints_to_grid (data/common.cc:205-231)
    cells = {}
    for i in range(n):
        cx = _cv_floor(f64(f32(x[i] - f32(grid.min_x))) * grid.inv_cell_width)
        cy = _cv_floor(f64(f32(y[i] - f32(grid.min_y))) * grid.inv_cell_height)
        if 0 <= cx < grid.num_cols and 0 <= cy < grid.num_rows:
            cells.setdefault((cx, cy), []).append(i)
    R = np.asarray(tgt["rot_cw"], f64).reshape(3, 3)
    t = np.asarray(tgt["trans_cw"], f64).reshape(3)
    c = np.asarray(tgt["cam_center"], f64).reshape(3)
    bits_k = np.unpackbits(tgt["desc"], axis=1)
    bits_l = np.unpackbits(lms["desc"], axis=1)
    best_out = np.full(m, -1, np.int32)
    dist_out = np.full(m, 0xFFFF, np.uint16)
    num_levels = len(scale_factors)
    for i in range(m):
        if lms.get("valid") is not None and not lms["valid"][i]:
            continue
        if tgt.get("skip") is not None and tgt["skip"][i]:
            continue
        X = np.asarray(lms["pos_w"][i], f64)
        pc = [R[r, 0] * X[0] + R[r, 1] * X[1] + R[r, 2] * X[2] + t[r] for r in range(3)]
        if pc[2] <= 0.0:
            continue
        z_inv = f64(1.0) / pc[2]
        u = cam.fx * pc[0] * z_inv + cam.cx
        v = cam.fy * pc[1] * z_inv + cam.cy
        q_xr = f32(u - cam.focal_x_baseline * z_inv)
        if not (cam.min_x < u < cam.max_x and cam.min_y < v < cam.max_y):
            continue
        d = X - c
        dist = np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        if dist < f64(lms["min_valid_dist"][i]) or f64(lms["max_valid_dist"][i]) < dist:
            continue
        nm = np.asarray(lms["obs_mean_normal"][i], f64)
        if d[0] * nm[0] + d[1] * nm[1] + d[2] * nm[2] < 0.5 * dist:
            continue
        ratio = f32(f32(lms["max_valid_dist_raw"][i]) / f32(dist))
        pred = int(np.ceil(f32(_logf(ratio) / f32(log_scale_factor))))
        pred = 0 if pred < 0 else (num_levels - 1 if pred >= num_levels else pred)
        ref_x, ref_y, r = f32(u), f32(v), f32(f32(margin) * f32(scale_factors[pred]))
        min_cx = max(0, _cv_floor(f64(f32(f32(ref_x - f32(grid.min_x)) - r)) * grid.inv_cell_width))
        max_cx = min(grid.num_cols - 1, _cv_ceil(f64(f32(f32(ref_x - f32(grid.min_x)) + r)) * grid.inv_cell_width))
        min_cy = max(0, _cv_floor(f64(f32(f32(ref_y - f32(grid.min_y)) - r)) * grid.inv_cell_height))
        max_cy = min(grid.num_rows - 1, _cv_ceil(f64(f32(f32(ref_y - f32(grid.min_y)) + r)) * grid.inv_cell_height))
        if grid.num_cols <= min_cx or max_cx < 0 or grid.num_rows <= min_cy or max_cy < 0:
            continue
        best_d, best_i = 256, -1
        for cx in range(min_cx, max_cx + 1):
            for cy in range(min_cy, max_cy + 1):
                for k in cells.get((cx, cy), []):
                    if not (abs(f32(x[k] - ref_x)) < r and abs(f32(y[k] - ref_y)) < r):
                        continue
                    lvl = int(octv[k])
                    if mode == 0:
                        if lvl < pred - 1 or pred < lvl:
                            continue
                    else:
                        if pred == 0:          # unsigned pred - 1 wraps: every candidate is rejected
                            continue
                        if lvl < pred - 1 or pred < lvl:
                            continue
                        e_x, e_y = u - f64(x[k]), v - f64(y[k])
                        kxr = f32(xr_all[k]) if xr_all is not None else f32(-1)
                        if kxr >= 0:
                            e_r = f32(q_xr - kxr)
                            err = e_x * e_x + e_y * e_y + f64(f32(e_r * e_r))
                            if f64(f32(7.81473)) < err * f64(f32(inv_level_sigma_sq[lvl])):
                                continue
                        else:
                            err = e_x * e_x + e_y * e_y
                            if f64(f32(5.99146)) < err * f64(f32(inv_level_sigma_sq[lvl])):
                                continue
                    hd = int((bits_l[i] != bits_k[k]).sum())
                    if hd < best_d:
                        best_d, best_i = hd, k
        if 50 < best_d:
            continue
        best_out[i], dist_out[i] = best_i, best_d
    return best_out, dist_out


def fuse_search_lines_python(cam, scale_factors_lsd, inv_level_sigma_sq_lsd, log_scale_factor_lsd, tgt, lms, margin):
    """match/fuse.cc:304-503 written from the reference text (second implementation beside the C++ oracle)."""
    f32, f64 = np.float32, np.float64
    sx, sy, ex, ey = (tgt[k].astype(f32) for k in ("sx", "sy", "ex", "ey"))
    octv = tgt["octave"]
    n, m = len(sx), len(lms["min_valid_dist"])
    R = np.asarray(tgt["rot_cw"], f64).reshape(3, 3)
    t = np.asarray(tgt["trans_cw"], f64).reshape(3)
    c = np.asarray(tgt["cam_center"], f64).reshape(3)
    bits_k = np.unpackbits(tgt["desc"], axis=1)
    bits_l = np.unpackbits(lms["desc"], axis=1)
    num_levels = len(scale_factors_lsd)
    best_out = np.full(m, -1, np.int32)
    dist_out = np.full(m, 0xFFFF, np.uint16)

    def reproject(X):
        pc = [R[r, 0] * X[0] + R[r, 1] * X[1] + R[r, 2] * X[2] + t[r] for r in range(3)]
        if pc[2] <= 0.0:
            return False, f64(0.0), f64(0.0)       # the reference leaves the reprojection unset; oracle rule: (0, 0)
        z_inv = f64(1.0) / pc[2]
        u = cam.fx * pc[0] * z_inv + cam.cx
        v = cam.fy * pc[1] * z_inv + cam.cy
        return bool(cam.min_x < u < cam.max_x and cam.min_y < v < cam.max_y), u, v

    def norm3(d):
        return np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
    for i in range(m):
        if lms.get("valid") is not None and not lms["valid"][i]:
            continue
        if tgt.get("skip") is not None and tgt["skip"][i]:
            continue
        P = np.asarray(lms["pos_w"][i], f64).reshape(6)
        S, E = P[:3], P[3:]
        in_s, su, sv = reproject(S)
        in_e, eu, ev = reproject(E)
        if not in_s and not in_e:
            continue
        M = np.array([0.5 * (S[k] + E[k]) for k in range(3)], f64)
        if not in_s or not in_e:
            if not reproject(M)[0]:
                continue
        ds, de = norm3(S - c), norm3(E - c)
        mn, mx = f64(lms["min_valid_dist"][i]), f64(lms["max_valid_dist"][i])
        if ds < mn or mx < ds or de < mn or mx < de:
            continue
        dm = norm3(M - c)
        ratio = f32(f32(lms["max_valid_dist_raw"][i]) / f32(dm))
        pred = int(np.ceil(f32(_logf(ratio) / f32(log_scale_factor_lsd))))
        pred = 0 if pred < 0 else (num_levels - 1 if pred >= num_levels else pred)
        r = f32(f32(margin) * f32(scale_factors_lsd[pred]))
        # get_keylines_in_cell: the line through the FLOAT reprojections (data/common.cc:324-327)
        ax, ay, bx, by = (f64(f32(w)) for w in (su, sv, eu, ev))
        f0, f1, f2 = ay * 1.0 - 1.0 * by, 1.0 * bx - ax * 1.0, ax * by - ay * bx
        l0, l1, l2 = sv * 1.0 - 1.0 * ev, 1.0 * eu - su * 1.0, su * ev - sv * eu
        best_d, best_i = 256, -1
        with np.errstate(divide="ignore", invalid="ignore"):
            fden, lden = np.sqrt(f0 * f0 + f1 * f1), np.sqrt(l0 * l0 + l1 * l1)
            for k in range(n):
                dsp = f32((f64(sx[k]) * f0 + f64(sy[k]) * f1 + f2) / fden)
                dep = f32((f64(ex[k]) * f0 + f64(ey[k]) * f1 + f2) / fden)
                if abs(dsp) > r or abs(dep) > r:
                    continue
                e_sp = (f64(sx[k]) * l0 + f64(sy[k]) * l1 + l2) / lden
                e_ep = (f64(ex[k]) * l0 + f64(ey[k]) * l1 + l2) / lden
                if f64(f32(5.99146)) < (e_sp * e_sp + e_ep * e_ep) * f64(f32(inv_level_sigma_sq_lsd[int(octv[k])])):
                    continue
                hd = int((bits_l[i] != bits_k[k]).sum())
                if hd < best_d:
                    best_d, best_i = hd, k
        if 50 < best_d:
            continue
        best_out[i], dist_out[i] = best_i, best_d
    return best_out, dist_out
