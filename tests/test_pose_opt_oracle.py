"""Self-tests of the pose-optimiser oracle.  g2o is absent from this environment, so parity with the
reference's g2o path is UNPINNED; these tests check the restatement against first principles instead:
convergence to the ground-truth pose, outlier recall, analytic-vs-numeric Jacobians, the < 5 observations rule."""
import numpy as np
import pytest

import synth


def _cam(plp, stereo=False):
    return plp.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, synth.COLS, synth.ROWS,
                                bf=synth.BF if stereo else -1.0, setup_type=1 if stereo else 0)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("stereo", [False, True])
def test_converges_and_flags_outliers(orc, plp, seed, stereo):
    T_gt, T_init, pts, lines = synth.make_pose_opt_scene(seed, stereo=stereo)
    cam = _cam(plp, stereo)
    T, pout, lout, n_inl, iters = orc.pose_optimize(cam, T_init, pts, lines if seed % 2 else None)
    err0 = np.linalg.norm(T_init - T_gt)
    err1 = np.linalg.norm(T - T_gt)
    assert err1 < 0.15 * err0 and err1 < 5e-3
    assert n_inl == len(pts) - int(pout.sum())
    assert 0.70 * len(pts) < n_inl < 0.92 * len(pts)  # 15 % gross outliers + chi-square tail
    assert 4 <= iters <= 40


def test_fewer_than_five_observations_is_a_no_op(orc, plp):
    T_gt, T_init, pts, lines = synth.make_pose_opt_scene(3, n_pts=4, n_lines=0)
    T, pout, _, n_inl, iters = orc.pose_optimize(_cam(plp), T_init, pts)
    assert n_inl == 0 and iters == 0 and np.array_equal(T, T_init)  # pose_optimizer.cc:153-156


def test_noise_free_problem_reaches_the_optimum(orc, plp):
    """Without outliers and noise a couple of LM steps must land on the ground truth: validates the analytic
    Jacobian / oplus convention pair (perspective_pose_opt_edge.cc:76-101, shot_vertex.h:58-62)."""
    T_gt, T_init, pts, _ = synth.make_pose_opt_scene(9, n_lines=0, outlier_frac=0.0, pose_sigma=(0.002, 0.005))
    R, t = T_gt[:3, :3], T_gt[:3, 3]
    Xc = pts["pos_w"] @ R.T + t
    pts["obs_x"] = synth.FX * Xc[:, 0] / Xc[:, 2] + synth.CX
    pts["obs_y"] = synth.FY * Xc[:, 1] / Xc[:, 2] + synth.CY
    T, pout, _, n_inl, _ = orc.pose_optimize(_cam(plp), T_init, pts, None, num_trials=1, num_each_iter=4)
    assert np.linalg.norm(T - T_gt) < 1e-4 and pout.sum() == 0


def test_lines_alone_constrain_the_pose(orc, plp):
    """Line edges (numeric Jacobian path) pull the pose to the ground truth when the 5 point edges carry
    almost no information."""
    T_gt, T_init, pts, lines = synth.make_pose_opt_scene(4, n_pts=5, n_lines=300, outlier_frac=0.0,
                                                         pose_sigma=(0.01, 0.02))
    pts["inv_sigma_sq"] = 1e-9
    T, pout, lout, n_inl, iters = orc.pose_optimize(_cam(plp), T_init, pts, lines)
    assert np.linalg.norm(T - T_gt) < 0.3 * np.linalg.norm(T_init - T_gt)
