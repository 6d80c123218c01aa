"""Config 1: the run_tum_rgbd_slam-shaped sequential loop through the host API.  On the CPU the oracle sits
behind the backend interface (plumbing: the trajectory of calls and a tracked sequence); on the GPU the same loop runs
through libplpslam_b200.so and must reproduce the oracle's trajectory call for call."""
import numpy as np
import pytest

import scene
import synth
from seq_backends import OracleBackend


def _run(plp, backend, n_frames=9, seed=99):
    from plpslam_b200.sequence import SequentialTracker
    seq = scene.PlanarSequence(seed=seed, n_frames=n_frames, tex_scale=1.3)
    cam = plp.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, synth.COLS, synth.ROWS)
    trk = SequentialTracker(backend, synth.ROWS, synth.COLS, cam)
    trace = []
    res = trk.run(seq, trace)
    return seq, res, trace


def test_sequential_loop_with_oracle_backend(plp, orc):
    seq, res, trace = _run(plp, OracleBackend(orc))
    n = len(seq.frames)
    assert len(res["poses"]) == n and len(trace) == n - 1 and res["tracked"] == n - 1
    for t in range(1, n):
        rel = np.linalg.norm(res["poses"][t] - seq.poses[t]) / np.linalg.norm(seq.poses[t])
        assert rel < 1e-2, (t, rel)  # monocular drift on a planar scene; the GPU test compares with the oracle at 1e-6
    for tr in trace:   # frame_tracker.cc:73-77 threshold, local map adds matches, second pose-opt keeps most of them
        assert tr["motion_matches"] >= 20 and tr["local_queries"] > 0 and tr["inliers"] >= 100
    assert set(res["stage_ms"][1]) == {"extract", "motion_match", "pose_opt_1", "local_map_match", "pose_opt_2"}


@pytest.mark.gpu
def test_sequential_loop_gpu_equals_oracle(plp, ctx, orc):
    from plpslam_b200.sequence import GpuBackend
    seq, want, wtrace = _run(plp, OracleBackend(orc))
    be = GpuBackend(plp, ctx, synth.ROWS, synth.COLS)
    _, got, gtrace = _run(plp, be)
    be.close()
    assert got["tracked"] == want["tracked"]
    for g, w in zip(gtrace, wtrace):
        assert np.array_equal(g["matched"], w["matched"]) and np.array_equal(g["best"], w["best"]), g["t"]
        assert g["inliers"] == w["inliers"]
    for Tg, Tw in zip(got["poses"], want["poses"]):
        assert np.linalg.norm(Tg - Tw) / np.linalg.norm(Tw) < 1e-6
