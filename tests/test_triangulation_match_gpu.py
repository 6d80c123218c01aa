"""GPU parity: robust::match_for_triangulation through the C ABI vs the oracle -- bit-exact match indices."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("stereo", [False, True])
def test_match_for_triangulation(ctx, orc, seed, stereo):
    kf1, kf2, fv1, fv2, E, ep = synth.make_triangulation_scene(seed, n=1200 if seed % 2 else 500, stereo=stereo)
    sf = synth.scale_factors()
    for check in (True, False):
        o, on = orc.match_for_triangulation(kf1, kf2, fv1, fv2, E, ep, sf, check)
        g, gn = ctx.match_for_triangulation(kf1, kf2, fv1, fv2, E, ep, sf, check)
        assert np.array_equal(g, o)
        assert gn == on
        # the libm acos of the reference and the deterministic acos give the same matches on this data
        ol, onl = orc.match_for_triangulation(kf1, kf2, fv1, fv2, E, ep, sf, check, libm=1)
        assert np.array_equal(ol, o)
    assert on > 30
    # every match pairs landmark-free keypoints of the same node and no keyframe-2 keypoint is used twice
    used = o[o >= 0]
    assert len(np.unique(used)) == len(used)
    assert not kf1["has_landmark"][o >= 0].any() and not kf2["has_landmark"][used].any()


def test_disjoint_vocabulary_nodes(ctx, orc):
    kf1, kf2, fv1, fv2, E, ep = synth.make_triangulation_scene(3, n=300)
    fv2 = ((fv2[0] + 100000).astype(np.uint32), fv2[1], fv2[2])  # no common node
    g, gn = ctx.match_for_triangulation(kf1, kf2, fv1, fv2, E, ep, synth.scale_factors())
    assert gn == 0 and (g == -1).all()


def test_landmark_compute_descriptor_batch(ctx, orc):
    rng = np.random.default_rng(4)
    counts = np.concatenate([[0, 1, 2, 3, 33, 64, 200], rng.integers(1, 30, 4000)])
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    base = synth.rand_desc(rng, len(counts))
    descs = np.concatenate([synth.flip_bits(rng, np.repeat(base[i:i + 1], c, 0), rng.integers(0, 60, c)) if c else
                            np.zeros((0, 32), np.uint8) for i, c in enumerate(counts)])
    # exact duplicates -> ties on the median distance (the first observation must win)
    descs[offsets[5]:offsets[5] + 8] = descs[offsets[5]]
    g = ctx.landmark_compute_descriptor_batch(descs, offsets)
    o = orc.landmark_compute_descriptor_batch(descs, offsets)
    assert np.array_equal(g, o)
    assert g[0] == -1 and g[1] == 0
