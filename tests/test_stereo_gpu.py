"""GPU parity: match::stereo::compute through the C ABI vs the oracle (bit-exact: Hamming winner indices, the float
sub-pixel x_right and depth of every left keypoint, and the median-correlation rejection)."""
import numpy as np
import pytest

import oracle_api
import synth

pytestmark = pytest.mark.gpu
BF, BASELINE = 47.906, 0.11  # example/euroc/EuRoC_stereo.yaml:34 (bf), baseline = bf / fx


def _run(ctx, orc, plp, left, right):
    h, w = left.shape
    p = oracle_api.orb_params()
    el, er = plp.OrbExtractor(ctx, h, w), plp.OrbExtractor(ctx, h, w)
    kl, dl = el.extract(left)
    kr, dr = er.extract(right)
    rl, rr = orc.orb_extract(p, left), orc.orb_extract(p, right)
    assert np.array_equal(kl, rl["kps"]) and np.array_equal(kr, rr["kps"])
    gx, gd, gb = el.stereo_compute(er, kl, dl, kr, dr, BF, BASELINE)
    ox, od, ob = orc.stereo_compute(rl, rr, el.scale_factors, el.inv_scale_factors, BF, BASELINE)
    assert np.array_equal(gb, ob), "Hamming-closest right keypoint"
    assert np.array_equal(gx, ox, equal_nan=True), "stereo_x_right"
    assert np.array_equal(gd, od, equal_nan=True), "depths"
    el.close()
    er.close()
    return kl, gx, gd


@pytest.mark.parametrize("seed,shape", [(3, (480, 752)), (8, (480, 640))])
def test_stereo_matches_oracle(ctx, orc, plp, seed, shape):
    left, right, disp = synth.make_stereo_pair(seed, *shape)
    kl, xr, dp = _run(ctx, orc, plp, left, right)
    ok = xr >= 0
    assert ok.sum() > 300
    # the recovered disparity is the rendered one (sanity of the test data, not a parity statement)
    d_true = disp[np.clip(kl["y"].astype(int), 0, shape[0] - 1), 0]
    err = np.abs((kl["x"] - xr)[ok] - d_true[ok])
    assert np.median(err) < 0.5


def test_stereo_edge_cases(ctx, orc, plp):
    # identical images: (near) zero disparity everywhere; best_disp <= 0 is replaced by 0.01 (stereo.cc:103-108)
    left = synth.make_texture(21, 480, 640)
    kl, xr, dp = _run(ctx, orc, plp, left, left.copy())
    ok = xr >= 0
    assert ok.sum() > 300 and np.median(dp[ok]) > 50.0  # tiny positive disparities -> large depths
    # unrelated right image: hardly any match survives the Hamming / correlation gates, same answer as the oracle
    _run(ctx, orc, plp, left, synth.make_texture(22, 480, 640))
    # flat right image: no right keypoint at all
    kl, xr, dp = _run(ctx, orc, plp, left, np.full((480, 640), 100, np.uint8))
    assert (xr == -1).all() and (dp == -1).all()
