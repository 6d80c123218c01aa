"""GPU parity: projection::match_frame_and_keyframe[_line] (relocalisation matchers) through the C ABI vs the oracle.
The oracle restates the whole reference function (visibility / distance gates, predict_scale_level, window match) and
emits the flattened queries the reference-side adapter hands to the C ABI; the GPU result must be bit-exact."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _kf_from_last(rng, last, Tc):
    m = len(last["pos_w"])
    P = np.asarray(last["pos_w"], np.float64).reshape(m, -1)
    mid = P[:, :3] if P.shape[1] == 3 else 0.5 * (P[:, :3] + P[:, 3:])
    c = -Tc[:3, :3].T @ Tc[:3, 3]
    dist = np.linalg.norm(mid - c, axis=1)
    # valid-distance interval around the true distance; some landmarks fail the 0.7 / 1.3 gates
    lo = (dist * rng.uniform(0.3, 1.3, m)).astype(np.float32)
    hi = (np.maximum(lo, dist) * rng.uniform(1.0, 2.5, m)).astype(np.float32)
    return dict(pos_w=P, min_valid_dist=lo, max_valid_dist=hi, desc=last["desc"], valid=last["valid"],
                angle=last.get("angle"))


@pytest.mark.parametrize("seed", range(8))
def test_match_frame_and_keyframe(ctx, orc, plp, seed):
    curr, last, Tc, _ = synth.make_tracking_scene(seed + 40, n_last=1000 if seed % 2 else 400)
    rng = np.random.default_rng(seed)
    kf = _kf_from_last(rng, last, Tc)
    grid = plp.capi.make_grid(synth.COLS, synth.ROWS)
    cam = plp.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, synth.COLS, synth.ROWS)
    sf = synth.scale_factors()
    # tracking_module / relocalizer call sites: (margin, hamm_dist_thr) = (10, 100), (3, 64) (module/relocalizer.cc:150-200)
    for margin, thr, check in [(10.0, 100, True), (3.0, 64, True), (20.0, 50, False), (10.0, 0, True)]:
        o, on, q = orc.match_frame_and_keyframe(grid, sf, cam, curr, Tc, kf, margin, thr, check)
        g, gn = ctx.match_frame_and_keyframe(grid, sf, curr, q, margin, thr, check)
        assert np.array_equal(g, o)
        assert gn == on
        if thr >= 64:
            assert on > 20
    assert 0 < q["valid"].sum() < len(q["valid"])           # the gates reject some landmarks
    assert len(np.unique(q["scale_level"][q["valid"] > 0])) > 2  # several predicted levels occur


@pytest.mark.parametrize("seed", range(6))
def test_match_frame_and_keyframe_line(ctx, orc, plp, seed):
    curr, last, Tc, _ = synth.make_line_scene(seed + 11, n_last=300)
    rng = np.random.default_rng(seed + 5)
    kf = _kf_from_last(rng, last, Tc)
    cam = plp.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, synth.COLS, synth.ROWS)
    sf_lsd = np.array([1.0, 2.0], np.float32)   # line_extractor.cc:54-55: _scale_factor = 2 (one level in practice)
    lsf = float(np.log(np.float32(2.0)))
    for margin, thr in [(10.0, 100), (5.0, 64), (30.0, 30)]:
        o, on, q = orc.match_frame_and_keyframe_line(sf_lsd[:1], lsf, cam, curr, Tc, kf, margin, thr)
        g, gn = ctx.match_frame_and_keyframe_line(sf_lsd[:1], curr, q, margin, thr)
        assert np.array_equal(g, o)
        assert gn == on
    assert on > 5


def test_empty_inputs(ctx, plp):
    grid = plp.capi.make_grid(synth.COLS, synth.ROWS)
    sf = synth.scale_factors()
    empty = dict(x=np.zeros(0, np.float32), y=np.zeros(0, np.float32), octave=np.zeros(0, np.int32),
                 desc=np.zeros((0, 32), np.uint8), angle=np.zeros(0, np.float32))
    q = dict(reproj_x=np.zeros(3, np.float32), reproj_y=np.zeros(3, np.float32), scale_level=np.zeros(3, np.int32),
             desc=np.zeros((3, 32), np.uint8), angle=np.zeros(3, np.float32), valid=np.ones(3, np.uint8))
    g, gn = ctx.match_frame_and_keyframe(grid, sf, empty, q, 10.0, 100, True)
    assert len(g) == 0 and gn == 0
