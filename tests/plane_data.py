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
