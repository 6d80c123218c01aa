"""The native thread-pool front-end chain of the oracle (oracle/frontend_mt.cc: bench.py's CPU arm) equals the chain
composed from the single-function restatements the parity tests use, for any thread count."""
import numpy as np

import oracle_api
import scene
import synth


def _py_chain(orc, pkg_capi, p, grid, cam, img, last, Tp, Tl):
    sf, isig = synth.scale_factors(), synth.inv_level_sigma_sq()
    r = orc.orb_extract(p, img)
    k = r["kps"]
    curr = dict(x=k["x"], y=k["y"], octave=k["octave"], angle=k["angle"], desc=r["desc"])
    m, nm = orc.match_current_and_last_frames(grid, sf, cam, curr, Tp, Tl, last, 20.0, True)
    if nm < 20:
        m, nm = orc.match_current_and_last_frames(grid, sf, cam, curr, Tp, Tl, last, 40.0, True)
    if nm < 20:
        return Tp, 0, 0, len(k)
    idx = np.nonzero(m >= 0)[0]
    pts = np.zeros(len(idx), oracle_api.PT_OBS_DTYPE)
    pts["pos_w"] = last["pos_w"][m[idx]]
    pts["obs_x"], pts["obs_y"] = k["x"][idx], k["y"][idx]
    pts["x_right"] = -1.0
    pts["inv_sigma_sq"] = isig[k["octave"][idx]]
    T, pout, _, n_inl, _ = orc.pose_optimize(cam, Tp, pts)
    return T, n_inl, int((pout == 0).sum()), len(k)


def test_native_chain_equals_composed_chain():
    import importlib.util
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    spec = importlib.util.spec_from_file_location("plp_capi_only", root / "structure-plp-slam_b200" / "capi.py")
    capi = importlib.util.module_from_spec(spec)
    sys.modules["plp_capi_only"] = capi
    spec.loader.exec_module(capi)
    orc = oracle_api.Oracle()
    p = oracle_api.orb_params()
    grid = capi.make_grid(synth.COLS, synth.ROWS)
    cam = capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, synth.COLS, synth.ROWS)
    seq = scene.PlanarSequence(seed=77, n_frames=4, tex_scale=1.2)
    rng = np.random.default_rng(3)
    lasts, preds, plast, imgs = [], [], [], []
    for t in range(1, 4):
        r = orc.orb_extract(p, seq.frames[t - 1])
        lasts.append(seq.last_frame_landmarks(t - 1, r["kps"], r["desc"]))
        preds.append(seq.predicted_pose(t, rng))
        plast.append(seq.poses[t - 1])
        imgs.append(seq.frames[t])
    imgs = np.stack(imgs)
    want = [_py_chain(orc, capi, p, grid, cam, imgs[b], lasts[b], preds[b], plast[b]) for b in range(3)]
    for threads in (1, 3):
        got = orc.frontend_track_batch(p, grid, cam, imgs, lasts, np.stack(preds), np.stack(plast), 20.0, threads)
        for b in range(3):
            assert np.array_equal(got["pose"][b], want[b][0])
            assert got["n_inliers"][b] == want[b][1] and got["num_valid"][b] == want[b][2] and got["n_kp"][b] == want[b][3]
        assert got["n_inliers"].min() >= 20
    assert orc.host_cpus() >= 1


def test_native_stereo_chain_equals_composed_chain():
    """orc_stereo_frontend_batch_mt (bench.py's stereo CPU arm) = orb_extract x2 + stereo_compute + line_extract x2."""
    orc = oracle_api.Oracle()
    p = oracle_api.orb_params()
    pairs = [synth.make_stereo_pair(11 + i, 240, 376, plp=True) for i in range(3)]
    left, right = np.stack([pr[0] for pr in pairs]), np.stack([pr[1] for pr in pairs])
    bf, baseline = 47.906, 0.11
    for threads in (1, 3):
        got = orc.stereo_frontend_batch_mt(p, left, right, bf, baseline, threads)
        for b in range(3):
            a, c = orc.orb_extract(p, left[b]), orc.orb_extract(p, right[b])
            _, dp, _ = orc.stereo_compute(a, c, synth.scale_factors(), 1.0 / synth.scale_factors(), bf, baseline)
            assert tuple(got["n_kp"][b]) == (len(a["kps"]), len(c["kps"]))
            assert got["n_stereo"][b] == int((dp > 0).sum())
            assert got["n_lines"][b, 0] == len(orc.line_extract(left[b])[0])
            assert got["n_lines"][b, 1] == len(orc.line_extract(right[b])[0])
