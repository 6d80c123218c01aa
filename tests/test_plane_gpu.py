"""GPU parity: Planar_Mapping_module plane RANSAC through the C ABI vs the oracle (both compile the same planemath.h text
without FMA contraction: equation, error, inlier flags and status must be bit-identical).

Verified on a B200 since round 1 (GPUTEST_r01: all cases passed)."""
import numpy as np
import pytest

import plane_data

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(5))
def test_plane_ransac_estimate(ctx, orc, seed):
    n = [300, 120, 800, 60, 18][seed]
    pts, valid, _, _ = plane_data.make_plane_cloud(seed + 10, n=n)
    smp = plane_data.draw_plane_samples(seed, valid, 50, 18)
    want = orc.plane_ransac(pts, valid, smp, plane_data.CFG_ESTIMATE)
    got = ctx.plane_ransac(pts, valid, smp, plane_data.CFG_ESTIMATE)
    assert got[0] == want[0]
    assert np.array_equal(got[1], want[1]) and got[2] == want[2]
    assert np.array_equal(got[3], want[3])


@pytest.mark.parametrize("seed", range(3))
def test_plane_ransac_update(ctx, orc, seed):
    pts, valid, _, _ = plane_data.make_plane_cloud(seed + 20, n=250, outlier_frac=0.1)
    eq0, err0 = orc.plane_fit(pts, np.nonzero(valid)[0][:40].astype(np.int32))
    smp = plane_data.draw_plane_samples(seed + 1, valid, 20, int(np.ceil(0.8 * len(pts))))
    want = orc.plane_ransac(pts, valid, smp, plane_data.CFG_UPDATE, eq0, err0)
    got = ctx.plane_ransac(pts, valid, smp, plane_data.CFG_UPDATE, eq0, err0)
    assert got[0] == want[0] and np.array_equal(got[1], want[1]) and got[2] == want[2] and np.array_equal(got[3], want[3])


def test_plane_ransac_edge_cases(ctx, orc, plp):
    pts, valid, _, _ = plane_data.make_plane_cloud(3, n=60)
    smp = plane_data.draw_plane_samples(0, valid, 10, 18)
    assert ctx.plane_ransac(pts[:10], None, smp % 10, plane_data.CFG_ESTIMATE)[0] == 0
    assert ctx.plane_ransac(pts[:10], None, smp % 10, plane_data.CFG_UPDATE)[0] == 2
    cloud = np.random.default_rng(1).uniform(-1, 1, (80, 3))
    s2 = plane_data.draw_plane_samples(2, np.ones(80), 15, 18)
    want, got = orc.plane_ransac(cloud, None, s2, plane_data.CFG_ESTIMATE), ctx.plane_ransac(cloud, None, s2, plane_data.CFG_ESTIMATE)
    assert got[0] == want[0] == 0 and np.array_equal(got[1], want[1]) and got[2] == want[2]
    with pytest.raises(plp.PlpError):
        ctx.plane_ransac(pts, valid, np.full((2, 18), 60, np.int32), plane_data.CFG_ESTIMATE)
