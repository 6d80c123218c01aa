"""GPU parity: match::fuse::{replace,detect}_duplication[_line] search through the C ABI vs the oracle (bit-exact
indices and distances), plus the adapter protocol with the GPU as the search backend vs the sequential reference loop."""
import numpy as np
import pytest

import fuse_data
import synth

pytestmark = pytest.mark.gpu


def _cam(plp, stereo=False):
    return plp.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, synth.COLS, synth.ROWS,
                                bf=synth.BF if stereo else -1.0, setup_type=1 if stereo else 0)


@pytest.mark.parametrize("seed", range(6))
def test_fuse_search_points(ctx, orc, plp, seed):
    stereo = bool(seed & 1)
    lms, targets = fuse_data.make_point_fuse_scene(seed, m=900 if seed % 3 else 300, num_targets=1 + seed % 4,
                                                   stereo=stereo)
    grid = plp.capi.make_grid(synth.COLS, synth.ROWS)
    cam = _cam(plp, stereo)
    sf, isg = synth.scale_factors(), fuse_data.inv_level_sigma_sq()
    matched = 0
    # call sites: replace_duplication margin 3 (fuse.h:66), detect_duplication margin 4 (global_optimization_module.cc:632)
    for margin, mode in [(3.0, 1), (4.0, 0), (10.0, 1), (0.5, 0)]:
        g_idx, g_dist = ctx.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, targets, lms, margin, mode)
        for t, tgt in enumerate(targets):
            o_idx, o_dist, _ = orc.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, tgt, lms, margin, mode)
            assert np.array_equal(g_idx[t], o_idx)
            assert np.array_equal(g_dist[t], o_dist)
            matched += (o_idx >= 0).sum()
    assert matched > 100


@pytest.mark.parametrize("seed", range(4))
def test_fuse_search_lines(ctx, orc, plp, seed):
    levels = 1 + 2 * (seed & 1)
    lms, targets = fuse_data.make_line_fuse_scene(seed + 3, m=300, num_targets=1 + seed % 3, num_levels=levels)
    cam = _cam(plp)
    sf = np.float32([1.0, 2.0, 4.0])[:levels]
    isg = (1.0 / (sf * sf)).astype(np.float32)
    lsf = float(np.log(np.float32(2.0)).astype(np.float32))
    matched = 0
    for margin in (10.0, 3.0, 40.0):   # mapping_module.cc:759 uses 10
        g_idx, g_dist = ctx.fuse_search_lines(cam, sf, isg, lsf, targets, lms, margin)
        for t, tgt in enumerate(targets):
            o_idx, o_dist, _ = orc.fuse_search_lines(cam, sf, isg, lsf, tgt, lms, margin)
            assert np.array_equal(g_idx[t], o_idx)
            assert np.array_equal(g_dist[t], o_dist)
            matched += (o_idx >= 0).sum()
    assert matched > 50


def test_fuse_full_size_batch(ctx, orc, plp):
    """mapping_module.cc:711-749 at config size: 20 target keyframes x 1000 landmarks, then 1 keyframe x 6000."""
    grid = plp.capi.make_grid(synth.COLS, synth.ROWS)
    cam = _cam(plp)
    sf, isg = synth.scale_factors(), fuse_data.inv_level_sigma_sq()
    lms, targets = fuse_data.make_point_fuse_scene(77, m=1000, num_targets=20)
    g_idx, g_dist = ctx.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, targets, lms, 3.0, 1)
    for t in (0, 7, 19):
        o_idx, o_dist, _ = orc.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, targets[t], lms, 3.0, 1)
        assert np.array_equal(g_idx[t], o_idx) and np.array_equal(g_dist[t], o_dist)
    lms, targets = fuse_data.make_point_fuse_scene(78, m=6000, num_targets=1, n_extra=400, obs_frac=0.2)
    g_idx, g_dist = ctx.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, targets, lms, 3.0, 1)
    o_idx, o_dist, _ = orc.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, targets[0], lms, 3.0, 1)
    assert np.array_equal(g_idx[0], o_idx) and np.array_equal(g_dist[0], o_dist)
    # a landmark matched in the batch is matched identically when searched alone (state-free search)
    pick = np.nonzero(o_idx >= 0)[0][:50]
    sub = {k: np.asarray(v)[pick] for k, v in lms.items()}
    tgt = dict(targets[0])
    tgt["skip"] = targets[0]["skip"][pick]
    s_idx, _ = ctx.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, [tgt], sub, 3.0, 1)
    assert np.array_equal(s_idx[0], o_idx[pick])


def test_fuse_edge_cases(ctx, orc, plp):
    grid = plp.capi.make_grid(synth.COLS, synth.ROWS)
    cam = _cam(plp)
    sf, isg = synth.scale_factors(), fuse_data.inv_level_sigma_sq()
    lms, targets = fuse_data.make_point_fuse_scene(5, m=64, num_targets=2)
    # an empty keyframe among the targets, landmarks behind the camera, all-invalid landmarks
    empty = dict(x=np.zeros(0, np.float32), y=np.zeros(0, np.float32), octave=np.zeros(0, np.int32),
                 desc=np.zeros((0, 32), np.uint8), rot_cw=np.eye(3), trans_cw=np.zeros(3), cam_center=np.zeros(3))
    lms["pos_w"][:8, 2] = -3.0
    g_idx, g_dist = ctx.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, [targets[0], empty, targets[1]], lms, 3.0, 1)
    assert np.all(g_idx[1] == -1) and np.all(g_dist[1] == 0xFFFF)
    for t, tgt in ((0, targets[0]), (2, targets[1])):
        o_idx, o_dist, _ = orc.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, tgt, lms, 3.0, 1)
        assert np.array_equal(g_idx[t], o_idx) and np.array_equal(g_dist[t], o_dist)
        assert np.all(o_idx[:8] == -1)
    lms["valid"][:] = 0
    g_idx, _ = ctx.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, targets, lms, 3.0, 1)
    assert np.all(g_idx == -1)
    # no landmarks: nothing to do, PLP_OK
    none = {k: np.asarray(v)[:0] for k, v in lms.items()}
    g_idx, _ = ctx.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, targets, none, 3.0, 1)
    assert g_idx.shape == (2, 0)
    # capacity is reported, not silently truncated
    big = dict(empty)
    big.update(x=np.zeros(4000, np.float32), y=np.zeros(4000, np.float32), octave=np.zeros(4000, np.int32),
               desc=np.zeros((4000, 32), np.uint8))
    lms["valid"][:] = 1
    with pytest.raises(plp.PlpError):
        ctx.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, [big], lms, 3.0, 1)


@pytest.mark.parametrize("seed", range(2))
def test_adapter_protocol_on_gpu(ctx, orc, plp, seed):
    """Batched GPU search + ordered effects + re-search == the reference's sequential loop (oracle search)."""
    import test_fuse_oracle as model_helpers
    lms, targets = fuse_data.make_point_fuse_scene(seed + 20, m=300, num_targets=4, n_extra=80)
    for t in targets:
        t.pop("skip")
    grid = plp.capi.make_grid(synth.COLS, synth.ROWS)
    cam = _cam(plp)
    sf, isg = synth.scale_factors(), fuse_data.inv_level_sigma_sq()
    lm_ids = [i for i in range(300) if lms["valid"][i]]

    def sub_lms(js, descs):
        ids = [lm_ids[j] for j in js]
        return {k: (np.asarray(v)[ids] if k != "desc" else np.asarray(descs)) for k, v in lms.items()}

    def gpu_batch(kfs, js, descs):
        return ctx.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, [targets[k] for k in kfs], sub_lms(js, descs),
                                      3.0, 1)[0]

    def orc_one(kf, js, descs):
        return orc.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, targets[kf], sub_lms(js, descs), 3.0, 1)[0]

    a = model_helpers._model_for(lms, targets, np.random.default_rng(seed))
    b = model_helpers._model_for(lms, targets, np.random.default_rng(seed))
    na = fuse_data.fuse_sequential(a, lm_ids, list(range(4)), orc_one)
    nb, researches = fuse_data.fuse_batched(b, lm_ids, list(range(4)), gpu_batch)
    assert na == nb and researches > 0
    sa, sb = a.state(), b.state()
    assert all(np.array_equal(x, y) for x, y in zip(sa[0], sb[0])) and sa[1] == sb[1]
    assert np.array_equal(sa[2], sb[2]) and np.array_equal(sa[3], sb[3])
