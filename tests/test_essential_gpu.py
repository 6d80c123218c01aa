"""GPU parity: solve::essential_solver::find_via_ransac through the C ABI vs the oracle.  Both sides compile the same
essmath.h text without FMA contraction, so the essential matrix, the score and the inlier flags must be bit-identical."""
import numpy as np
import pytest

import ess_data

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(6))
def test_essential_ransac(ctx, orc, seed):
    n = [400, 60, 1000, 150, 8, 33][seed]
    b1, b2, matches, _ = ess_data.make_two_view(seed + 10, n=n, outlier_frac=0.3 if n > 8 else 0.0)
    samples = ess_data.draw_samples(seed, len(matches), 50)      # robust.cc:232: find_via_ransac(50, false)
    for recompute in (False, True):
        want = orc.essential_ransac(b1, b2, matches, samples, recompute)
        got = ctx.essential_ransac(b1, b2, matches, samples, recompute)
        assert got[0] == want[0]
        assert np.array_equal(got[1], want[1])
        assert np.array_equal(got[2], want[2])          # bit-identical doubles
        assert got[3] == want[3]
    assert want[0] == 1 or n <= 33


def test_essential_ransac_edge_cases(ctx, orc, plp):
    b1, b2, matches, _ = ess_data.make_two_view(3, n=40)
    # fewer than 8 matches: invalid, nothing computed (essential_solver.cc:45-49)
    got = ctx.essential_ransac(b1, b2, matches[:7], np.zeros((5, 8), np.int32), False)
    assert got[0] == 0 and got[1].sum() == 0
    # no iterations: best_score_ stays 0 -> invalid
    got = ctx.essential_ransac(b1, b2, matches, np.zeros((0, 8), np.int32), False)
    assert got[0] == 0 and got[1].sum() == 0 and got[3] == 0.0
    # degenerate sample (the same match eight times): same answer as the oracle, no crash
    samples = np.zeros((4, 8), np.int32)
    samples[1] = np.arange(8)
    want = orc.essential_ransac(b1, b2, matches, samples, False)
    got = ctx.essential_ransac(b1, b2, matches, samples, False)
    assert got[0] == want[0] and np.array_equal(got[1], want[1]) and got[3] == want[3]
    # out-of-range indices are rejected
    bad = matches.copy()
    bad[0, 0] = 10**6
    with pytest.raises(plp.PlpError):
        ctx.essential_ransac(b1, b2, bad, samples, False)
    with pytest.raises(plp.PlpError):
        ctx.essential_ransac(b1, b2, matches, np.full((1, 8), 40, np.int32), False)
