"""Seeded synthetic inputs shared by the parity tests and bench.py (no reference data needed)."""
from __future__ import annotations

import numpy as np

FX, FY, CX, CY = 535.4, 539.2, 320.1, 247.6  # TUM fr3 intrinsics (example/tum_rgbd/TUM_RGBD_mono_3.yaml)
COLS, ROWS = 640, 480
BF = 47.906


def scale_factors(num_levels=8, sf=1.2):
    out = np.ones(num_levels, np.float32)
    for i in range(1, num_levels):
        out[i] = np.float32(sf) * out[i - 1]
    return out


def rand_desc(rng, n):
    return rng.integers(0, 256, size=(n, 32), dtype=np.uint8)


def flip_bits(rng, desc, nbits):
    """Flip `nbits` random bits (per row) of 32-byte descriptors."""
    out = desc.copy()
    n = desc.shape[0]
    for i in range(n):
        k = int(nbits[i]) if np.ndim(nbits) else int(nbits)
        if k <= 0:
            continue
        pos = rng.choice(256, size=k, replace=False)
        for p in pos:
            out[i, p >> 3] ^= np.uint8(1 << (p & 7))
    return out


def so3_exp(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * (K @ K)


def make_pose(rng, rot_sigma=0.02, trans_sigma=0.05):
    T = np.eye(4)
    T[:3, :3] = so3_exp(rng.normal(0, rot_sigma, 3))
    T[:3, 3] = rng.normal(0, trans_sigma, 3)
    return T


def project(T, X, fx=FX, fy=FY, cx=CX, cy=CY):
    Xc = X @ T[:3, :3].T + T[:3, 3]
    z = Xc[:, 2]
    return np.stack([fx * Xc[:, 0] / z + cx, fy * Xc[:, 1] / z + cy], 1), z


def make_tracking_scene(seed, n_last=1000, n_extra=300, stereo=False, num_levels=8, dup_frac=0.15):
    """A last frame with landmarks and a current frame that re-observes most of them.

    Returns (curr, last, Tc, Tl): dicts in the layout of Context.match_current_and_last_frames.
    Keypoint coordinates are level-grid values (integer * scale factor) like real ORB output, and a
    fraction of current keypoints are near-duplicates so that claims/ties actually occur.
    """
    rng = np.random.default_rng(seed)
    sf = scale_factors(num_levels)
    Tl = np.eye(4)
    Tc = make_pose(rng, 0.01, 0.03)
    X = np.stack([rng.uniform(-4, 4, n_last), rng.uniform(-3, 3, n_last), rng.uniform(2, 12, n_last)], 1)
    last_oct = rng.integers(0, num_levels, n_last).astype(np.int32)
    last_desc = rand_desc(rng, n_last)
    last_angle = rng.uniform(0, 360, n_last).astype(np.float32)
    last_valid = (rng.random(n_last) > 0.1).astype(np.uint8)

    uv, z = project(Tc, X)
    # current keypoints: re-observations (with pixel noise, on the level grid) + duplicates + clutter
    oct_c = np.clip(last_oct + rng.integers(-1, 2, n_last), 0, num_levels - 1).astype(np.int32)
    noise = rng.normal(0, 1.5, (n_last, 2)) * sf[oct_c][:, None]
    pts = uv + noise
    s = sf[oct_c][:, None].astype(np.float32)
    pts = (np.round(pts / s) * s).astype(np.float32)
    desc_c = flip_bits(rng, last_desc, rng.integers(5, 60, n_last))
    ang_c = (last_angle + rng.normal(0, 4, n_last)).astype(np.float32) % np.float32(360)
    # a few wildly rotated ones to exercise the orientation histogram
    wild = rng.random(n_last) < 0.08
    ang_c[wild] = rng.uniform(0, 360, wild.sum()).astype(np.float32)
    keep = rng.random(n_last) > 0.15
    xs, ys, octs, descs, angs = [pts[keep, 0]], [pts[keep, 1]], [oct_c[keep]], [desc_c[keep]], [ang_c[keep]]
    # duplicates: same cell, same or near-equal descriptors (ties on distance!)
    nd = int(dup_frac * keep.sum())
    src = rng.choice(np.nonzero(keep)[0], nd, replace=False)
    dpts = pts[src] + (rng.integers(-2, 3, (nd, 2)) * sf[oct_c[src]][:, None]).astype(np.float32)
    ddesc = desc_c[src].copy()
    mod = rng.random(nd) < 0.5
    ddesc[mod] = flip_bits(rng, ddesc[mod], rng.integers(1, 4, mod.sum()))
    xs.append(dpts[:, 0]); ys.append(dpts[:, 1]); octs.append(oct_c[src]); descs.append(ddesc); angs.append(ang_c[src])
    # clutter
    xs.append(rng.uniform(-5, COLS + 5, n_extra).astype(np.float32))
    ys.append(rng.uniform(-5, ROWS + 5, n_extra).astype(np.float32))
    octs.append(rng.integers(0, num_levels, n_extra).astype(np.int32))
    descs.append(rand_desc(rng, n_extra))
    angs.append(rng.uniform(0, 360, n_extra).astype(np.float32))
    x = np.concatenate(xs).astype(np.float32)
    y = np.concatenate(ys).astype(np.float32)
    perm = rng.permutation(len(x))
    curr = dict(x=x[perm], y=y[perm], octave=np.concatenate(octs)[perm].astype(np.int32),
                desc=np.concatenate(descs)[perm], angle=np.concatenate(angs)[perm].astype(np.float32))
    n = len(x)
    curr["claimed"] = (rng.random(n) < 0.05).astype(np.uint8)
    if stereo:
        xr = np.full(n, -1.0, np.float32)
        has = rng.random(n) < 0.7
        xr[has] = (curr["x"][has] - rng.uniform(2, 40, has.sum())).astype(np.float32)
        curr["x_right"] = xr
    last = dict(pos_w=X, octave=last_oct, angle=last_angle, desc=last_desc, valid=last_valid)
    return curr, last, Tc, Tl


def make_landmark_queries(seed, curr, m=1500, num_levels=8):
    """Local-map queries for match_frame_and_landmarks built around existing keypoints."""
    rng = np.random.default_rng(seed + 7919)
    sf = scale_factors(num_levels)
    n = len(curr["x"])
    src = rng.integers(0, n, m)
    lvl = np.clip(curr["octave"][src] + rng.integers(0, 2, m), 0, num_levels - 1).astype(np.int32)
    rx = (curr["x"][src] + rng.normal(0, 2.0, m) * sf[lvl]).astype(np.float32)
    ry = (curr["y"][src] + rng.normal(0, 2.0, m) * sf[lvl]).astype(np.float32)
    desc = flip_bits(rng, curr["desc"][src], rng.integers(0, 70, m))
    q = dict(reproj_x=rx, reproj_y=ry, scale_level=lvl, desc=desc, valid=(rng.random(m) > 0.1).astype(np.uint8))
    if "x_right" in curr:
        q["x_right"] = (rx - rng.uniform(2, 40, m)).astype(np.float32)
    return q


def make_line_scene(seed, n_last=200, n_extra=60, num_levels=1):
    """Keylines for the *_line matchers."""
    rng = np.random.default_rng(seed)
    Tl = np.eye(4)
    Tc = make_pose(rng, 0.01, 0.03)
    P0 = np.stack([rng.uniform(-4, 4, n_last), rng.uniform(-3, 3, n_last), rng.uniform(2, 10, n_last)], 1)
    d = rng.normal(0, 1, (n_last, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    P1 = P0 + d * rng.uniform(0.5, 3.0, (n_last, 1))
    # push a few endpoints behind the camera / out of the image
    far = rng.random(n_last) < 0.1
    P1[far, 2] -= 15
    pos_w = np.concatenate([P0, P1], 1)
    last_desc = rand_desc(rng, n_last)
    last = dict(pos_w=pos_w, octave=np.zeros(n_last, np.int32), desc=last_desc,
                valid=(rng.random(n_last) > 0.1).astype(np.uint8))
    a, _ = project(Tc, P0)
    b, _ = project(Tc, P1)
    keep = rng.random(n_last) > 0.2
    sx = a[keep, 0] + rng.normal(0, 1.0, keep.sum())
    sy = a[keep, 1] + rng.normal(0, 1.0, keep.sum())
    ex = b[keep, 0] + rng.normal(0, 1.0, keep.sum())
    ey = b[keep, 1] + rng.normal(0, 1.0, keep.sum())
    desc = flip_bits(rng, last_desc[keep], rng.integers(5, 60, keep.sum()))
    # duplicates
    nd = keep.sum() // 5
    src = rng.integers(0, keep.sum(), nd)
    sx = np.concatenate([sx, sx[src] + rng.normal(0, 0.5, nd), rng.uniform(0, COLS, n_extra)])
    sy = np.concatenate([sy, sy[src] + rng.normal(0, 0.5, nd), rng.uniform(0, ROWS, n_extra)])
    ex = np.concatenate([ex, ex[src] + rng.normal(0, 0.5, nd), rng.uniform(0, COLS, n_extra)])
    ey = np.concatenate([ey, ey[src] + rng.normal(0, 0.5, nd), rng.uniform(0, ROWS, n_extra)])
    desc = np.concatenate([desc, desc[src], rand_desc(rng, n_extra)])
    n = len(sx)
    perm = rng.permutation(n)
    curr = dict(sx=sx[perm].astype(np.float32), sy=sy[perm].astype(np.float32), ex=ex[perm].astype(np.float32),
                ey=ey[perm].astype(np.float32), octave=np.zeros(n, np.int32), desc=desc[perm],
                claimed=(rng.random(n) < 0.05).astype(np.uint8),
                ratio_level=rng.integers(0, 3, n).astype(np.int32))
    return curr, last, Tc, Tl


def make_line_queries(seed, curr, m=300):
    rng = np.random.default_rng(seed + 104729)
    n = len(curr["sx"])
    src = rng.integers(0, n, m)
    q = dict(sp_x=(curr["sx"][src] + rng.normal(0, 1.5, m)).astype(np.float32),
             sp_y=(curr["sy"][src] + rng.normal(0, 1.5, m)).astype(np.float32),
             ep_x=(curr["ex"][src] + rng.normal(0, 1.5, m)).astype(np.float32),
             ep_y=(curr["ey"][src] + rng.normal(0, 1.5, m)).astype(np.float32),
             scale_level=np.zeros(m, np.int32), desc=flip_bits(rng, curr["desc"][src], rng.integers(0, 70, m)),
             valid=(rng.random(m) > 0.1).astype(np.uint8))
    return q


# ------------------------------------------------------------------------------------- images
def make_texture(seed, h=480, w=640, n_rect=400, n_blob=2000):
    """Seeded textured scene: random-contrast rectangles (axis-aligned + rotated) and Gaussian blobs over
    1/f noise: > 1000 FAST corners per level budget, line-rich."""
    import cv2
    rng = np.random.default_rng(seed)
    # 1/f noise
    f = np.fft.fft2(rng.normal(0, 1, (h, w)))
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.fftfreq(w)[None, :]
    rad = np.sqrt(fx * fx + fy * fy)
    rad[0, 0] = 1.0
    img = np.real(np.fft.ifft2(f / rad))
    img = (img - img.min()) / (img.max() - img.min()) * 90.0 + 80.0
    img = img.astype(np.float32)
    sc = max(h, w) / 640.0
    for _ in range(n_rect):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        rw, rh = rng.uniform(8, 90) * sc, rng.uniform(8, 90) * sc
        ang = rng.uniform(0, 180) if rng.random() < 0.5 else 0.0
        box = cv2.boxPoints(((float(cx), float(cy)), (float(rw), float(rh)), float(ang)))
        val = float(rng.uniform(0, 255))
        a = float(rng.uniform(0.4, 1.0))
        layer = img.copy()
        cv2.fillPoly(layer, [np.round(box).astype(np.int32)], val)
        img = (1 - a) * img + a * layer
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for _ in range(n_blob):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        s = rng.uniform(1.0, 3.5) * sc
        amp = rng.uniform(-90, 90)
        x0, x1 = int(max(0, cx - 4 * s)), int(min(w, cx + 4 * s + 1))
        y0, y1 = int(max(0, cy - 4 * s)), int(min(h, cy + 4 * s + 1))
        if x1 <= x0 or y1 <= y0:
            continue
        g = np.exp(-((xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2) / (2 * s * s))
        img[y0:y1, x0:x1] += (amp * g).astype(np.float32)
    return np.clip(np.round(img), 0, 255).astype(np.uint8)


def make_toy_corner_image(variant=1):
    """The toy images of test/PLPSLAM/feature/orb_extractor.cc:27-83 (white image, one anti-aliased black
    rectangle).  Returns (image, corner_xy)."""
    import cv2
    if variant == 1:  # extract_toy_sample_1
        img = np.full((600, 600), 255, np.uint8)
        cv2.rectangle(img, (300, 300), (600, 600), 0, -1, cv2.LINE_AA)
        return img, (300, 300)
    img = np.full((2000, 2000), 255, np.uint8)  # extract_toy_sample_2
    cv2.rectangle(img, (0, 0), (1800, 1800), 0, -1, cv2.LINE_AA)
    return img, (1800, 1800)


# ------------------------------------------------------------------------------------- pose optimiser
def inv_level_sigma_sq(num_levels=8, sf=1.2):
    out = np.ones(num_levels, np.float32)
    s = np.float32(1.0)
    for i in range(1, num_levels):
        s = np.float32(sf) * s
        out[i] = np.float32(1.0) / (s * s)
    return out


def plucker_from_endpoints(P, Q):
    """Line::set_pos_in_world (data/landmark_line.cc:60-77): (q x p, p - q)."""
    return np.concatenate([np.cross(Q, P), P - Q], axis=-1)


def make_pose_opt_scene(seed, n_pts=1000, n_lines=200, outlier_frac=0.15, stereo=False, pose_sigma=(0.02, 0.05)):
    """points in front of the camera, octave-dependent pixel noise, gross
    outliers, perturbed initial pose.  Returns (T_gt, T_init, pts, lines) in the C-ABI layouts."""
    rng = np.random.default_rng(seed)
    T_gt = make_pose(rng, 0.2, 0.5)
    sf = scale_factors()
    isig = inv_level_sigma_sq()
    Xc = np.stack([rng.uniform(-5, 5, n_pts), rng.uniform(-5, 5, n_pts), rng.uniform(2, 12, n_pts)], 1)
    R, t = T_gt[:3, :3], T_gt[:3, 3]
    Xw = (Xc - t) @ R  # R^T (Xc - t)
    octave = rng.choice(8, n_pts, p=np.array([217, 181, 151, 126, 105, 87, 73, 60]) / 1000.0)
    uv = np.stack([FX * Xc[:, 0] / Xc[:, 2] + CX, FY * Xc[:, 1] / Xc[:, 2] + CY], 1)
    uv += rng.normal(0, 1.0, (n_pts, 2)) * sf[octave][:, None]
    out = rng.random(n_pts) < outlier_frac
    uv[out] = np.stack([rng.uniform(0, COLS, out.sum()), rng.uniform(0, ROWS, out.sum())], 1)
    pts = np.zeros(n_pts, np.dtype([("pos_w", "<f8", 3), ("obs_x", "<f4"), ("obs_y", "<f4"), ("x_right", "<f4"),
                                    ("inv_sigma_sq", "<f4")]))
    pts["pos_w"] = Xw
    pts["obs_x"], pts["obs_y"] = uv[:, 0], uv[:, 1]
    pts["x_right"] = -1.0
    if stereo:
        has = rng.random(n_pts) < 0.7
        xr = uv[:, 0] - BF / Xc[:, 2] + rng.normal(0, 1.0, n_pts) * sf[octave]
        pts["x_right"][has] = xr[has]
    pts["inv_sigma_sq"] = isig[octave]
    lines = np.zeros(n_lines, np.dtype([("plucker", "<f8", 6), ("sp_x", "<f4"), ("sp_y", "<f4"), ("ep_x", "<f4"),
                                        ("ep_y", "<f4"), ("inv_sigma_sq", "<f4"), ("pad", "<f4")]))
    if n_lines:
        Pc = np.stack([rng.uniform(-4, 4, n_lines), rng.uniform(-3, 3, n_lines), rng.uniform(3, 10, n_lines)], 1)
        d = rng.normal(0, 1, (n_lines, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        d[:, 2] *= 0.3
        Qc = Pc + d * rng.uniform(0.5, 3.0, (n_lines, 1))
        Pw, Qw = (Pc - t) @ R, (Qc - t) @ R
        lines["plucker"] = plucker_from_endpoints(Pw, Qw)
        sp = np.stack([FX * Pc[:, 0] / Pc[:, 2] + CX, FY * Pc[:, 1] / Pc[:, 2] + CY], 1) + rng.normal(0, 1.0, (n_lines, 2))
        ep = np.stack([FX * Qc[:, 0] / Qc[:, 2] + CX, FY * Qc[:, 1] / Qc[:, 2] + CY], 1) + rng.normal(0, 1.0, (n_lines, 2))
        lo = rng.random(n_lines) < outlier_frac
        sp[lo] += rng.normal(0, 40, (lo.sum(), 2))
        lines["sp_x"], lines["sp_y"], lines["ep_x"], lines["ep_y"] = sp[:, 0], sp[:, 1], ep[:, 0], ep[:, 1]
        lines["inv_sigma_sq"] = 1.0
    xi = np.concatenate([rng.normal(0, pose_sigma[0], 3), rng.normal(0, pose_sigma[1], 3)])
    dT = np.eye(4)
    dT[:3, :3] = so3_exp(xi[:3])
    dT[:3, 3] = xi[3:]
    T_init = dT @ T_gt
    return T_gt, T_init, pts, lines


def make_line_image(seed, h=480, w=640, n_patch=44, noise=2.0):
    """Seeded line-rich scene (the benchmark images must be line-rich to reach the ~200 line
    features per frame of BASELINE.json): rotated patches of parallel stripes (facades, shelves, floor boards) with
    random grey levels painted over a smooth ramp, anti-aliased, plus sensor noise."""
    import cv2
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = 110.0 + 40.0 * np.sin(xx / w * 2.1 + rng.uniform(0, 3)) + 30.0 * np.cos(yy / h * 1.7 + rng.uniform(0, 3))
    img = img.astype(np.float32)
    sc = max(h, w) / 640.0
    for _ in range(n_patch):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        length = rng.uniform(90, 260) * sc
        n_str = int(rng.integers(3, 8))
        widths = rng.uniform(9, 24, n_str) * sc
        ang = rng.uniform(0, np.pi)
        ca, sa = np.cos(ang), np.sin(ang)
        off = -widths.sum() / 2
        base = float(rng.uniform(20, 235))
        for k in range(n_str):
            val = float(np.clip(base + (1 if k % 2 else -1) * rng.uniform(35, 110), 5, 250))
            u0, u1 = -length / 2, length / 2
            v0, v1 = off, off + widths[k]
            off = v1
            q = np.array([[u0, v0], [u1, v0], [u1, v1], [u0, v1]])
            pts = np.stack([cx + q[:, 0] * ca - q[:, 1] * sa, cy + q[:, 0] * sa + q[:, 1] * ca], 1)
            cv2.fillPoly(img, [np.round(pts * 16).astype(np.int32)], val, lineType=cv2.LINE_AA, shift=4)
    img = cv2.GaussianBlur(img, (0, 0), 0.8)
    img += rng.normal(0, noise, (h, w)).astype(np.float32)
    return np.clip(np.round(img), 0, 255).astype(np.uint8)


def make_stereo_pair(seed, h=480, w=752, bf=47.906, n_bands=6, plp=False):
    """Rectified stereo pair: the right image is the left texture shifted by a per-row-band
    disparity bf / depth (sub-pixel, bilinear) plus independent sensor noise."""
    import cv2
    rng = np.random.default_rng(seed)
    left = make_plp_texture(seed, h, w) if plp else make_texture(seed, h, w)
    depths = rng.uniform(1.5, 12.0, n_bands)
    right = np.zeros_like(left)
    edges = np.linspace(0, h, n_bands + 1).astype(int)
    xx = np.arange(w, dtype=np.float32)[None, :].repeat(h, 0)
    yy = np.arange(h, dtype=np.float32)[:, None].repeat(w, 1)
    disp = np.zeros((h, w), np.float32)
    for i in range(n_bands):
        disp[edges[i]:edges[i + 1]] = bf / depths[i]
    right = cv2.remap(left, xx + disp, yy, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    noise = rng.normal(0, 1.5, (h, w))
    right = np.clip(np.round(right.astype(np.float64) + noise), 0, 255).astype(np.uint8)
    return left, right, disp


def make_triangulation_scene(seed, n=1200, stereo=False, num_levels=8, n_nodes=90):
    """Two keyframes observing a common point cloud, their relative essential matrix, and synthetic BoW feature vectors
    (points near each other in descriptor space share a node; a few nodes exist in only one keyframe)."""
    rng = np.random.default_rng(seed)
    T1 = np.eye(4)
    T2 = make_pose(rng, 0.05, 0.4)
    X = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(2, 12, n)], 1)
    node_of_point = rng.integers(0, n_nodes, n)
    desc = rand_desc(rng, n)

    def view(T, keep_frac, noise_px):
        uv, z = project(T, X)
        keep = (rng.random(n) < keep_frac) & (z > 0.1)
        idx = np.nonzero(keep)[0]
        uv = uv[idx] + rng.normal(0, noise_px, (len(idx), 2))
        b = np.stack([(uv[:, 0] - CX) / FX, (uv[:, 1] - CY) / FY, np.ones(len(idx))], 1)
        b /= np.linalg.norm(b, axis=1, keepdims=True)
        d = flip_bits(rng, desc[idx], rng.integers(0, 45, len(idx)))
        nodes = node_of_point[idx].copy()
        # clutter: unrelated keypoints
        nc = len(idx) // 5
        bc = rng.normal(0, 1, (nc, 3)); bc[:, 2] = np.abs(bc[:, 2]) + 1.0
        bc /= np.linalg.norm(bc, axis=1, keepdims=True)
        b = np.concatenate([b, bc]); d = np.concatenate([d, rand_desc(rng, nc)])
        nodes = np.concatenate([nodes, rng.integers(0, n_nodes + 10, nc)])
        m = len(b)
        perm = rng.permutation(m)
        b, d, nodes = b[perm], d[perm], nodes[perm]
        f = dict(desc=d, bearings=b, angle=rng.uniform(0, 360, m).astype(np.float32),
                 octave=rng.integers(0, num_levels, m).astype(np.int32),
                 has_landmark=(rng.random(m) < 0.3).astype(np.uint8))
        if stereo:
            xr = np.full(m, -1.0, np.float32)
            has = rng.random(m) < 0.5
            xr[has] = rng.uniform(1, 600, has.sum()).astype(np.float32)
            f["x_right"] = xr
        # feature vector: ascending node ids, random order inside a node (DBoW2 appends in keypoint order; any order is legal)
        ids = np.unique(nodes)
        ids = ids[rng.random(len(ids)) < 0.9]
        offsets, indices = [0], []
        for nid in ids:
            members = np.nonzero(nodes == nid)[0]
            indices.extend(members.tolist())
            offsets.append(len(indices))
        fv = (ids.astype(np.uint32), np.array(offsets, np.int32), np.array(indices, np.uint32))
        return f, fv, idx

    kf1, fv1, _ = view(T1, 0.8, 0.4)
    kf2, fv2, _ = view(T2, 0.8, 0.4)
    # consistent angles for true pairs are irrelevant for parity; E_12 = [t_12]x R_12 with x1^T E x2 = 0
    T12 = T1 @ np.linalg.inv(T2)
    R12, t12 = T12[:3, :3], T12[:3, 3]
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    E12 = tx @ R12
    c1 = -T1[:3, :3].T @ T1[:3, 3]
    ep = T2[:3, :3] @ c1 + T2[:3, 3]
    ep = ep / np.linalg.norm(ep)
    return kf1, kf2, fv1, fv2, E12, ep


def make_plp_texture(seed, h=480, w=640, tile=(80, 106), cover=0.95, noise=2.0):
    """Point-AND-line-rich scene for the full PLP front end (BASELINE north_star: ~1000 ORB + ~200 line features per
    640x480 frame): the corner-rich texture of make_texture with non-overlapping tiles of parallel stripes (facades,
    shelves, floor boards) painted over `cover` of the tile grid.  LSD works at half resolution and keeps segments
    >= 60 px, so ~200 of them need about two thirds of the image as edges ~100 px long and >= 12 px apart: the stripes run
    along the long side of their tile (+- 0.1 rad), which makes every edge a full tile long."""
    import cv2
    rng = np.random.default_rng(seed + 7919)
    img = make_texture(seed, h, w).astype(np.float32)
    th, tw = tile
    ny, nx = max(1, h // th), max(1, w // tw)
    cells = [(j, i) for j in range(ny) for i in range(nx)]
    rng.shuffle(cells)
    for (j, i) in cells[:int(round(cover * len(cells)))]:
        y0, x0 = j * h // ny, i * w // nx
        y1, x1 = (j + 1) * h // ny, (i + 1) * w // nx
        hh, ww = y1 - y0, x1 - x0
        pad = 6
        yy, xx = np.mgrid[0:hh, 0:ww].astype(np.float32)
        ang = (np.pi / 2 if ww >= hh else 0.0) + rng.uniform(-0.1, 0.1)
        period = rng.uniform(12, 15)
        phase = rng.uniform(0, period)
        u = (xx - ww / 2) * np.cos(ang) + (yy - hh / 2) * np.sin(ang) + phase
        k = np.floor(u / period).astype(np.int64)
        levels = rng.uniform(15, 240, 64)
        levels[1::2] = np.clip(levels[::2] + rng.choice([-1, 1], 32) * rng.uniform(45, 110, 32), 5, 250)
        patch = levels[np.mod(k, 64)].astype(np.float32)
        m = np.zeros((hh, ww), np.float32)
        m[pad:hh - pad, pad:ww - pad] = 1.0
        m = cv2.GaussianBlur(m, (0, 0), 1.0)
        img[y0:y1, x0:x1] = img[y0:y1, x0:x1] * (1 - m) + patch * m
    img = cv2.GaussianBlur(img, (0, 0), 0.7)
    img += rng.normal(0, noise, (h, w)).astype(np.float32)
    return np.clip(np.round(img), 0, 255).astype(np.uint8)
