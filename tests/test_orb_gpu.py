"""GPU parity: ORB extraction through the C ABI vs the oracle -- keypoints, order, angles and descriptors
bit-exact, stage by stage (pyramid, FAST candidates, final output)."""
import numpy as np
import pytest

import oracle_api
import synth

pytestmark = pytest.mark.gpu


def _compare(ext, orc, p, img, mask=None):
    r = orc.orb_extract(p, img, mask=mask, debug=True)
    kps, desc = ext.extract(img, mask=mask)
    # stage 1: pyramid (public image_pyramid_)
    for l in range(p.num_levels):
        assert np.array_equal(ext.pyramid_level(0, l), r["pyramid"][l]), f"pyramid level {l}"
    # stage 2: FAST candidates per level, in the reference's cell / row-major order
    off = 0
    for l in range(p.num_levels):
        c = r["cands"][off: off + r["cands_per_level"][l]]
        off += r["cands_per_level"][l]
        g = ext.debug_candidates(0, l)
        assert len(g) == len(c), f"candidate count level {l}"
        for f in ("x", "y", "response"):
            assert np.array_equal(g[f], c[f]), f"candidates level {l} field {f}"
    # stage 3: final keypoints and descriptors
    assert len(kps) == len(r["kps"])
    for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(kps[f], r["kps"][f]), f
    assert np.array_equal(desc, r["desc"])
    return len(kps)


@pytest.mark.parametrize("seed,shape", [(1234, (480, 640)), (77, (480, 752)), (5, (480, 640)), (6, (376, 1241))])
def test_extract_matches_oracle(ctx, orc, plp, seed, shape):
    img = synth.make_texture(seed, shape[0], shape[1])
    p = oracle_api.orb_params()
    ext = plp.OrbExtractor(ctx, shape[0], shape[1])
    n = _compare(ext, orc, p, img)
    assert n > 900
    ext.close()


def test_extract_2000_keypoints_and_other_pyramids(ctx, orc, plp):
    img = synth.make_texture(11)
    for (mk, sf, lv, ini, mn) in [(2000, 1.2, 8, 20, 7), (500, 1.5, 4, 30, 10), (300, 1.2, 1, 20, 7)]:
        p = oracle_api.orb_params(mk, sf, lv, ini, mn)
        ext = plp.OrbExtractor(ctx, 480, 640, mk, sf, lv, ini, mn)
        _compare(ext, orc, p, img)
        ext.close()


def test_low_texture_and_noise_images(ctx, orc, plp):
    p = oracle_api.orb_params()
    ext = plp.OrbExtractor(ctx, 480, 640)
    flat = np.full((480, 640), 128, np.uint8)
    kps, desc = ext.extract(flat)
    assert len(kps) == 0 and desc.shape == (0, 32)
    rng = np.random.default_rng(3)
    # weak texture: the minimum-threshold fallback (orb_extractor.cc:407-412) fires in most cells
    weak = (128 + rng.normal(0, 4.0, (480, 640))).clip(0, 255).astype(np.uint8)
    _compare(ext, orc, p, weak)
    # pure noise: tens of thousands of candidates per level (global-memory path of the quadtree kernel)
    noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    _compare(ext, orc, p, noise)
    # half flat / half textured: empty cells next to busy ones
    half = synth.make_texture(21)
    half[:, :300] = 90
    _compare(ext, orc, p, half)
    ext.close()


def test_reference_toy_corner(ctx, orc, plp):
    # test/PLPSLAM/feature/orb_extractor.cc:27-54 through the CUDA path
    img, (cx, cy) = synth.make_toy_corner_image(1)
    ext = plp.OrbExtractor(ctx, 600, 600, 2000)
    kps, desc = ext.extract(img)
    assert len(kps) == len(desc) > 0 and desc.dtype == np.uint8
    for k in kps:
        tol = 2.0 * ext.scale_factors[k["octave"]]
        assert abs(k["x"] - cx) <= tol and abs(k["y"] - cy) <= tol
    p = oracle_api.orb_params(2000)
    _compare(ext, orc, p, img)
    ext.close()


def test_mask(ctx, orc, plp):
    # test/PLPSLAM/feature/orb_extractor.cc:125-359: keypoints stay out of the masked region
    img = synth.make_texture(31)
    mask = np.full(img.shape, 255, np.uint8)
    mask[:, :320] = 0
    mask[100:200, 400:500] = 0
    p = oracle_api.orb_params()
    ext = plp.OrbExtractor(ctx, 480, 640)
    _compare(ext, orc, p, img, mask=mask)
    kps, _ = ext.extract(img, mask=mask)
    assert np.all(kps["x"] >= 320 - 1e-3)
    ext.close()


def test_batch_equals_single(ctx, orc, plp):
    imgs = np.stack([synth.make_texture(100 + i) for i in range(5)])
    ext = plp.OrbExtractor(ctx, 480, 640, max_batch=8)
    res = ext.extract_batch(imgs)
    p = oracle_api.orb_params()
    for b in range(5):
        r = orc.orb_extract(p, imgs[b])
        assert len(res[b][0]) == len(r["kps"])
        for f in ("x", "y", "angle", "response", "octave"):
            assert np.array_equal(res[b][0][f], r["kps"][f])
        assert np.array_equal(res[b][1], r["desc"])
    ext.close()


def test_empty_image_is_silent(ctx, plp):
    ext = plp.OrbExtractor(ctx, 480, 640)
    kps, desc = ext.extract(None)  # orb_extractor.cc:76-79
    assert len(kps) == 0
    ext.close()
