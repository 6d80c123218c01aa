"""Seeded point clouds for the Planar_Mapping_module plane RANSAC tests."""
import numpy as np


def make_plane_cloud(seed, n=300, outlier_frac=0.25, noise=0.004):
    """`n` landmarks linked to one plane instance: most lie on a random plane (Gaussian noise), the rest are off-plane."""
    rng = np.random.default_rng(seed)
    nrm = rng.normal(0, 1, 3)
    nrm /= np.linalg.norm(nrm)
    d = rng.uniform(-3, 3)
    u = np.cross(nrm, [1.0, 0.2, 0.1])
    u /= np.linalg.norm(u)
    v = np.cross(nrm, u)
    uv = rng.uniform(-2, 2, (n, 2))
    pts = uv[:, :1] * u + uv[:, 1:] * v - d * nrm + rng.normal(0, noise, (n, 1)) * nrm
    out = rng.random(n) < outlier_frac
    pts[out] += rng.uniform(0.05, 0.6, (out.sum(), 1)) * nrm * rng.choice([-1, 1], (out.sum(), 1))
    valid = (rng.random(n) > 0.05).astype(np.uint8)
    return pts, valid, np.append(nrm, d), ~out


def draw_plane_samples(seed, valid, num_iter, sample_size):
    """The index draws of planar_mapping_module.cc:447-457 / :620-632: uniform indices, erased landmarks rejected
    (repeats allowed, as in the reference)."""
    rng = np.random.default_rng(seed)
    ok = np.nonzero(valid)[0]
    return rng.choice(ok, (num_iter, sample_size), replace=True).astype(np.int32)


CFG_ESTIMATE = dict(mode=0, points_per_ransac=18, planar_distance_thresh=0.02, final_error_thresh=0.002,
                    inliers_ratio_thr=0.6)
CFG_UPDATE = dict(mode=1, points_per_ransac=18, planar_distance_thresh=0.02, final_error_thresh=0.002,
                  inliers_ratio_thr=0.6, initial_best_error=0.01)


def plane_ransac_python(pos, valid, samples, cfg, eq0=(0, 0, 0, 0), err0=0.0):
    """planar_mapping_module.cc:412-733 written from the reference text with numpy's SVD as estimate_plane_SVD (a second
    implementation beside the C++ oracle; fits agree to ~1e-12, so statuses and inlier sets must agree on these scenes)."""
    pos = np.asarray(pos, np.float64).reshape(-1, 3)
    n = len(pos)
    valid = np.ones(n, bool) if valid is None else np.asarray(valid).astype(bool)
    P = cfg["points_per_ransac"]
    inl_out = np.zeros(n, np.uint8)
    eq, plane_err = np.array(eq0, np.float64), float(err0)

    def fit(idx):
        X = pos[idx]
        c = X.mean(0)
        U = np.linalg.svd((X - c).T, full_matrices=True)[0]
        nrm = U[:, 2] / np.linalg.norm(U[:, 2])
        d = -nrm @ c
        return np.append(nrm, d), abs(np.linalg.norm(X @ nrm + d) / len(idx))

    def dist(e, X):
        return np.abs((X @ e[:3] + e[3]) / np.linalg.norm(e[:3]))
    if n == 0:
        return 0, eq, plane_err, inl_out
    if n < P:
        return (2 if cfg["mode"] == 1 else 0), eq, plane_err, inl_out
    best_error = cfg.get("initial_best_error", 0.0) if cfg["mode"] == 1 else np.finfo(np.float64).max
    best_found, best_list = False, []
    for it in range(len(samples)):
        e, residual = fit(np.asarray(samples[it]))
        if residual < best_error:
            best_error = residual
        eq, plane_err = e, residual
        inliers = [j for j in range(n) if valid[j] and dist(eq, pos[j]) < cfg["planar_distance_thresh"]]
        if cfg["mode"] == 0:
            eligible = len(inliers) / n > cfg["inliers_ratio_thr"] and len(inliers) >= P
        else:
            eligible = len(inliers) >= P
        if eligible:
            e2, error = fit(np.asarray(inliers))
            if error < best_error:
                best_error, eq, plane_err, best_list, best_found = error, e2, error, inliers, True
                if cfg["mode"] == 0 and error < cfg["final_error_thresh"]:
                    break
    if not best_found or best_error > cfg["final_error_thresh"]:
        return 0, eq, plane_err, inl_out
    kept = [j for j in best_list if valid[j] and dist(eq, pos[j]) < cfg["planar_distance_thresh"]]
    if cfg["mode"] == 1 and len(kept) < P:
        return 2, eq, plane_err, inl_out
    inl_out[kept] = 1
    return 1, eq, plane_err, inl_out
