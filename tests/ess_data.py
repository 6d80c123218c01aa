"""Seeded two-view scenes for the solve::essential_solver tests."""
import numpy as np

import synth


def make_two_view(seed, n=400, outlier_frac=0.3, noise=2e-3):
    """Bearings of `n` world points in two cameras (shot 1, shot 2), a match list with gross outliers, and E_21 truth."""
    rng = np.random.default_rng(seed)
    X = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(3, 12, n)], 1)
    T1 = synth.make_pose(rng, 0.02, 0.05)
    T2 = synth.make_pose(rng, 0.08, 0.6)
    def bearings(T):
        Xc = X @ T[:3, :3].T + T[:3, 3]
        b = Xc / np.linalg.norm(Xc, axis=1, keepdims=True) + rng.normal(0, noise, Xc.shape)
        return b / np.linalg.norm(b, axis=1, keepdims=True)
    b1, b2 = bearings(T1), bearings(T2)
    p1, p2 = rng.permutation(n), rng.permutation(n)
    bb1, bb2 = b1[p1], b2[p2]                     # keypoint order of each shot
    inv1, inv2 = np.argsort(p1), np.argsort(p2)
    good = rng.random(n) > outlier_frac
    second = np.where(good, inv2[np.arange(n)], rng.integers(0, n, n))
    matches = np.stack([inv1[np.arange(n)], second], 1).astype(np.int32)[rng.permutation(n)]
    R21 = T2[:3, :3] @ T1[:3, :3].T
    t21 = -R21 @ T1[:3, 3] + T2[:3, 3]
    tx = np.array([[0, -t21[2], t21[1]], [t21[2], 0, -t21[0]], [-t21[1], t21[0], 0]])
    return bb1, bb2, matches, tx @ R21


def draw_samples(seed, num_matches, num_iter=50):
    """util::create_random_array(8, 0, num_matches - 1) per RANSAC iteration: 8 distinct indices, ascending."""
    rng = np.random.default_rng(seed)
    return np.stack([np.sort(rng.choice(num_matches, 8, replace=False)) for _ in range(num_iter)]).astype(np.int32)
