"""Pins the oracle against the reference's own known-answer tests.

Each test restates a gtest case from /root/reference/test/PLPSLAM (file:line cited) -- the
numbers are the reference's, the implementation under test is oracle/.
"""
import numpy as np
import pytest

import synth


# ---- test/PLPSLAM/match/base.cc:11-66
@pytest.mark.parametrize("a,b,expected", [(0b01010101, 0b01010101, 0), (0b01010101, 0b10101010, 256),
                                          (0b01100110, 0b00111100, 128)])
def test_hamming_known_answers(orc, a, b, expected):
    d1 = np.full(32, a, np.uint8)
    d2 = np.full(32, b, np.uint8)
    assert orc.hamming_32(d1, d2) == expected
    assert orc.hamming_64(d1, d2) == expected


def test_hamming_32_64_agree_random(orc):
    rng = np.random.default_rng(0)
    d = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    M = orc.hamming_matrix(d[:100], d[100:])
    ref = np.unpackbits(d[:100, None, :] ^ d[None, 100:, :], axis=2).sum(2)
    assert np.array_equal(M, ref)
    for i in range(0, 100, 7):
        assert orc.hamming_64(d[i], d[100 + i]) == ref[i, i]


# ---- test/PLPSLAM/match/angle_checker.cc:13-192
_DELTAS = [34.8, 34.9, 35.0, 35.1, 35.2, 323.8, 323.9, 324.1, 324.2, 126.9, 127.0, 127.1, 0.0, 90.0, 180.0, 270.0]
_MATCH = [35] * 5 + [324] * 4 + [127] * 3 + [0, 60, 180, 270]


@pytest.mark.parametrize("top,valid_set", [(1, {35}), (2, {35, 324}), (3, {35, 324, 127}),
                                           (30, {35, 324, 127, 0, 60, 180, 270})])
def test_angle_checker_partitions(orc, top, valid_set):
    valid = set(orc.angle_checker(_DELTAS, _MATCH, 30, top, valid=True).tolist())
    invalid = set(orc.angle_checker(_DELTAS, _MATCH, 30, top, valid=False).tolist())
    assert valid == valid_set
    assert invalid == set(_MATCH) - valid_set
    assert len(orc.angle_checker(_DELTAS, _MATCH, 30, top, valid=True)) + \
        len(orc.angle_checker(_DELTAS, _MATCH, 30, top, valid=False)) == len(_MATCH)


# ---- test/PLPSLAM/data/common_get_cell_indices.cc (undistorted camera cases)
def test_get_cell_indices_valid_and_invalid(orc, plp):
    cols, rows = 2000, 1000
    g = plp.capi.make_grid(cols, rows)
    eps = 0.01
    w, h = 1.0 / g.inv_cell_width, 1.0 / g.inv_cell_height
    for ix in range(g.num_cols):
        for iy in range(0, g.num_rows, 5):
            for (x, y) in [(ix * w + eps, iy * h + eps), ((ix + 1) * w - eps, iy * h + eps),
                           (ix * w + eps, (iy + 1) * h - eps), ((ix + 1) * w - eps, (iy + 1) * h - eps)]:
                ok, cx, cy = orc.get_cell_indices(g, x, y)
                assert ok and cx == ix and cy == iy
    # corners / centres
    assert orc.get_cell_indices(g, 0.0, 0.0) == (True, 0, 0)
    assert orc.get_cell_indices(g, cols - eps, rows - eps) == (True, g.num_cols - 1, g.num_rows - 1)
    # outside the bounds -> invalid
    for (x, y) in [(-eps, 10.0), (10.0, -eps), (cols + eps, 10.0), (10.0, rows + eps)]:
        assert not orc.get_cell_indices(g, x, y)[0]


def test_keypoints_in_cell_matches_bruteforce_window(orc, plp):
    """get_keypoints_in_cell == the window predicate, in (cell-x, cell-y, insertion) order."""
    rng = np.random.default_rng(3)
    g = plp.capi.make_grid(640, 480)
    n = 800
    x = rng.uniform(0, 640, n).astype(np.float32)
    y = rng.uniform(0, 480, n).astype(np.float32)
    octv = rng.integers(0, 8, n).astype(np.int32)
    for _ in range(50):
        rx, ry = np.float32(rng.uniform(-20, 660)), np.float32(rng.uniform(-20, 500))
        margin = np.float32(rng.uniform(3, 40))
        mn, mx = int(rng.integers(-1, 7)), int(rng.integers(-1, 8))
        got = orc.get_keypoints_in_cell(g, x, y, octv, rx, ry, margin, mn, mx)
        chk = (mn > 0) or (mx >= 0)
        sel = (np.abs(x - rx) < margin) & (np.abs(y - ry) < margin)
        if chk:
            sel &= octv >= mn
            if mx >= 0:
                sel &= octv <= mx
        cx = np.floor(x.astype(np.float64) * g.inv_cell_width).astype(int)
        cy = np.floor(y.astype(np.float64) * g.inv_cell_height).astype(int)
        idx = np.nonzero(sel)[0]
        order = np.lexsort((idx, cy[idx], cx[idx]))
        assert np.array_equal(got, idx[order])
