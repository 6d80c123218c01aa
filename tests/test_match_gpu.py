"""GPU parity: Hamming matchers through the C ABI vs the CPU oracle -- bit-exact indices."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def test_hamming_known_answers_gpu(ctx):
    # test/PLPSLAM/match/base.cc:11-66 through the CUDA path
    a = np.stack([np.full(32, 0b01010101, np.uint8), np.full(32, 0b01010101, np.uint8), np.full(32, 0b01100110, np.uint8)])
    b = np.stack([np.full(32, 0b01010101, np.uint8), np.full(32, 0b10101010, np.uint8), np.full(32, 0b00111100, np.uint8)])
    M = ctx.hamming_matrix(a, b)
    assert M[0, 0] == 0 and M[1, 1] == 256 and M[2, 2] == 128


@pytest.mark.parametrize("na,nb", [(1, 1), (33, 65), (1000, 1000), (257, 2049)])
def test_hamming_matrix_matches_oracle(ctx, orc, na, nb):
    rng = np.random.default_rng(na * 7 + nb)
    a, b = synth.rand_desc(rng, na), synth.rand_desc(rng, nb)
    assert np.array_equal(ctx.hamming_matrix(a, b), orc.hamming_matrix(a, b))


@pytest.mark.parametrize("nq,nt", [(1, 1), (200, 200), (333, 1027), (5, 0)])
def test_hamming_nn_matches_oracle(ctx, orc, nq, nt):
    rng = np.random.default_rng(nq + nt)
    q = synth.rand_desc(rng, nq)
    t = synth.rand_desc(rng, nt)
    if nt > 10:  # plant exact duplicates so that ties on distance occur
        t[nt // 2] = t[3]
        q[0] = t[3]
    gi, gd = ctx.hamming_nn(q, t)
    oi, od = orc.hamming_nn(q, t)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("stereo", [False, True])
def test_match_current_and_last_frames(ctx, orc, plp, seed, stereo):
    curr, last, Tc, Tl = synth.make_tracking_scene(seed, n_last=1000 if seed % 3 else 300, stereo=stereo)
    grid = plp.capi.make_grid(synth.COLS, synth.ROWS)
    cam = plp.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, synth.COLS, synth.ROWS,
                               bf=synth.BF if stereo else -1.0, setup_type=1 if stereo else 0)
    sf = synth.scale_factors()
    if stereo and seed % 2:  # force the forward/backward branches (projection.cc:231-238)
        Tc = Tc.copy()
        Tc[2, 3] += 0.5 if seed % 4 == 1 else -0.5
    for margin, check in [(20.0, True), (10.0, True), (40.0, False)]:
        g, gn = ctx.match_current_and_last_frames(grid, sf, cam, curr, Tc, Tl, last, margin, check)
        o, on = orc.match_current_and_last_frames(grid, sf, cam, curr, Tc, Tl, last, margin, check)
        assert np.array_equal(g, o)
        assert gn == on
        assert on > 20  # the scene really produces matches


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("stereo", [False, True])
def test_match_frame_and_landmarks(ctx, orc, plp, seed, stereo):
    curr, _, _, _ = synth.make_tracking_scene(seed + 100, n_last=900, stereo=stereo)
    q = synth.make_landmark_queries(seed, curr, m=1500 if seed % 2 else 4000)
    grid = plp.capi.make_grid(synth.COLS, synth.ROWS)
    sf = synth.scale_factors()
    for margin, ratio in [(5.0, 0.8), (10.0, 0.8), (20.0, 0.6)]:
        g, gn = ctx.match_frame_and_landmarks(grid, sf, curr, q, margin, ratio)
        o, on = orc.match_frame_and_landmarks(grid, sf, curr, q, margin, ratio)
        assert np.array_equal(g, o)
        assert gn == on and on > 20


def test_match_edge_cases(ctx, orc, plp):
    grid = plp.capi.make_grid(synth.COLS, synth.ROWS)
    cam = plp.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, synth.COLS, synth.ROWS)
    sf = synth.scale_factors()
    curr, last, Tc, Tl = synth.make_tracking_scene(5, n_last=50, n_extra=0)
    empty = {k: v[:0] for k, v in curr.items()}
    g, gn = ctx.match_current_and_last_frames(grid, sf, cam, empty, Tc, Tl, last, 20.0)
    assert len(g) == 0 and gn == 0
    last0 = {k: v[:0] for k, v in last.items()}
    g, gn = ctx.match_current_and_last_frames(grid, sf, cam, curr, Tc, Tl, last0, 20.0)
    assert np.all(g == -1) and gn == 0
    # every query invalid
    last_inv = dict(last, valid=np.zeros(len(last["octave"]), np.uint8))
    g, gn = ctx.match_current_and_last_frames(grid, sf, cam, curr, Tc, Tl, last_inv, 20.0)
    assert np.all(g == -1) and gn == 0
    # all identical descriptors at one location: the earliest query wins, the rest cascade
    n = 40
    same = dict(x=np.full(n, 100.0, np.float32) + np.arange(n, dtype=np.float32) * 0.0,
                y=np.full(n, 100.0, np.float32), octave=np.zeros(n, np.int32),
                desc=np.zeros((n, 32), np.uint8), angle=np.zeros(n, np.float32))
    q = dict(reproj_x=np.full(60, 100.0, np.float32), reproj_y=np.full(60, 100.0, np.float32),
             scale_level=np.zeros(60, np.int32), desc=np.zeros((60, 32), np.uint8))
    g, gn = ctx.match_frame_and_landmarks(grid, sf, same, q, 5.0, 2.0)
    o, on = orc.match_frame_and_landmarks(grid, sf, same, q, 5.0, 2.0)
    assert np.array_equal(g, o) and gn == on == n  # 40-deep dependency chain


@pytest.mark.parametrize("seed", range(8))
def test_brute_force_match(ctx, orc, seed):
    rng = np.random.default_rng(seed)
    n_kf, n_frm = (1000, 1000) if seed % 2 else (400, 1300)
    kf_desc = synth.rand_desc(rng, n_kf)
    src = rng.integers(0, n_kf, n_frm)
    frm_desc = synth.flip_bits(rng, kf_desc[src], rng.integers(0, 45, n_frm))
    # exact duplicates in the frame -> ties and claim conflicts
    frm_desc[n_frm // 2:n_frm // 2 + 50] = frm_desc[:50]
    kf_angle = rng.uniform(0, 360, n_kf).astype(np.float32)
    frm_angle = ((kf_angle[src] + rng.normal(0, 5, n_frm)) % 360).astype(np.float32)
    frm_angle[rng.random(n_frm) < 0.1] = 13.0
    kf_valid = (rng.random(n_kf) > 0.2).astype(np.uint8)
    for ratio, check in [(0.8, False), (0.8, True), (0.6, True)]:
        g, gn = ctx.brute_force_match(frm_desc, frm_angle, kf_desc, kf_angle, kf_valid, ratio, check)
        o, on = orc.brute_force_match(frm_desc, frm_angle, kf_desc, kf_angle, kf_valid, ratio, check)
        assert np.array_equal(g, o) and gn == on and on > 10


@pytest.mark.parametrize("seed", range(8))
def test_line_matchers(ctx, orc, plp, seed):
    curr, last, Tc, Tl = synth.make_line_scene(seed)
    cam = plp.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, synth.COLS, synth.ROWS)
    sf = np.array([1.0, 2.0], np.float32)
    for margin in (10.0, 20.0):
        g, gn = ctx.match_current_and_last_frames_line(sf, cam, curr, Tc, Tl, last, margin)
        o, on = orc.match_current_and_last_frames_line(sf, cam, curr, Tc, Tl, last, margin)
        assert np.array_equal(g, o) and gn == on and on > 5
    q = synth.make_line_queries(seed, curr)
    for margin, ratio in [(5.0, 0.8), (15.0, 0.6)]:
        g, gn = ctx.match_frame_and_landmarks_line(sf, curr, q, margin, ratio)
        o, on = orc.match_frame_and_landmarks_line(sf, curr, q, margin, ratio)
        assert np.array_equal(g, o) and gn == on and on > 5
