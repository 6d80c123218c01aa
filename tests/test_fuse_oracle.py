"""CPU tests of the match::fuse restatement (oracle), of the host-derived predict_scale_level threshold table of the
product library (no device work) and of the adapter protocol (batched search + ordered effects + re-search of landmarks
whose descriptor was recomputed) against the reference's landmark-by-landmark loop on a map-state model."""
import numpy as np
import pytest

import fuse_data
import synth


def _cam(plp, stereo=False):
    return plp.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, synth.COLS, synth.ROWS,
                                bf=synth.BF if stereo else -1.0, setup_type=1 if stereo else 0)


def test_level_thresholds_reproduce_host_logf(plp, orc):
    """count(ratio >= thr[k]) == landmark::predict_scale_level for random ratios and for every float next to a threshold."""
    rng = np.random.default_rng(0)
    for sf, levels in [(1.2, 8), (2.0, 1), (2.0, 3), (1.1, 12), (1.4142135, 5)]:
        lsf = float(np.log(np.float32(sf)).astype(np.float32))
        thr = plp.capi.fuse_level_thresholds(lsf, levels)
        assert np.all(np.diff(thr[1:]) > 0)
        ratios = np.exp(rng.uniform(-1.0, np.log(sf) * (levels + 1), 4000)).astype(np.float32)
        near = []
        for k in range(1, levels):
            t = thr[k]
            f = t
            for _ in range(40):
                f = np.nextafter(f, np.float32(0))
            for _ in range(80):
                near.append(f)
                f = np.nextafter(f, np.float32(np.inf))
        ratios = np.concatenate([ratios, np.asarray(near, np.float32), np.float32([0.0, 1.0, 1e-30, 1e30])])
        for r in ratios:
            want = orc.predict_scale_level(float(r), 1.0, lsf, levels)   # ratio = max_valid_dist / 1
            got = int(sum(r >= thr[k] for k in range(1, levels)))
            assert got == want, (sf, levels, r)


def test_level_thresholds_reject_bad_arguments(plp):
    with pytest.raises(plp.PlpError):
        plp.capi.fuse_level_thresholds(0.0, 8)
    with pytest.raises(plp.PlpError):
        plp.capi.fuse_level_thresholds(0.18, 64)


@pytest.mark.parametrize("seed", range(3))
def test_fuse_points_oracle_properties(plp, orc, seed):
    lms, targets = fuse_data.make_point_fuse_scene(seed, m=500, num_targets=2, stereo=bool(seed & 1))
    grid = plp.capi.make_grid(synth.COLS, synth.ROWS)
    cam = _cam(plp, bool(seed & 1))
    sf, isg = synth.scale_factors(), fuse_data.inv_level_sigma_sq()
    for tgt in targets:
        rep, rd, lvl = orc.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, tgt, lms, 3.0, 1)
        det, dd, _ = orc.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, tgt, lms, 3.0, 0)
        assert (rep >= 0).sum() > 40 and (det >= 0).sum() >= (rep >= 0).sum() // 2
        # gates: invalid / skipped landmarks never match; matched distances <= HAMMING_DIST_THR_LOW
        assert np.all(rep[(lms["valid"] == 0) | (tgt["skip"] != 0)] == -1)
        assert np.all(rd[rep >= 0] <= 50) and np.all(rd[rep < 0] == 0xFFFF)
        # the unsigned level gate of replace_duplication (fuse.cc:232): predicted level 0 rejects every candidate
        assert np.all(rep[lvl == 0] == -1)
        # the matched keypoint satisfies the window and the level gate, and its distance is what the oracle reports
        for i in np.nonzero(rep >= 0)[0][:60]:
            j = rep[i]
            assert lvl[i] - 1 <= tgt["octave"][j] <= lvl[i]
            assert orc.hamming_32(lms["desc"][i], tgt["desc"][j]) == rd[i]
        # detect mode has no chi-square gate: whenever replace matched, detect matched something at least as close
        both = (rep >= 0) & (det >= 0)
        assert np.all(dd[both] <= rd[both])
        assert len(np.unique(lvl[lvl >= 0])) >= 4


@pytest.mark.parametrize("seed", range(3))
def test_fuse_lines_oracle_properties(plp, orc, seed):
    lms, targets = fuse_data.make_line_fuse_scene(seed + 3, num_levels=1 + 2 * (seed & 1))
    cam = _cam(plp)
    levels = 1 + 2 * (seed & 1)
    sf = np.float32([1.0, 2.0, 4.0])[:levels]
    isg = (1.0 / (sf * sf)).astype(np.float32)
    lsf = float(np.log(np.float32(2.0)).astype(np.float32))
    total = 0
    for tgt in targets:
        best, dist, lvl = orc.fuse_search_lines(cam, sf, isg, lsf, tgt, lms, 10.0)
        total += (best >= 0).sum()
        assert np.all(best[(lms["valid"] == 0) | (tgt["skip"] != 0)] == -1)
        assert np.all(dist[best >= 0] <= 50)
        for i in np.nonzero(best >= 0)[0][:40]:
            assert orc.hamming_32(lms["desc"][i], tgt["desc"][best[i]]) == dist[i]
    assert total > 30


def _model_for(lms, targets, rng):
    """Map state: every target keyframe already holds landmarks (ids >= m) on a third of its keypoints, some shared
    between keyframes, so that fuse finds duplicates and replace() fires."""
    m = len(lms["desc"])
    kf_descs = [t["desc"] for t in targets]
    extra = 0
    observations = []
    per_kf = []
    for k, t in enumerate(targets):
        n = len(t["x"])
        idx = rng.choice(n, n // 3, replace=False)
        per_kf.append(idx)
        for i in idx:
            observations.append((m + extra, k, int(i)))
            extra += 1
    # give some of the existing landmarks a second / third observation in other keyframes (more observations than lm)
    lm_desc = np.concatenate([lms["desc"], synth.rand_desc(rng, extra)])
    model = fuse_data.MapModel(kf_descs, lm_desc, observations)
    # the checked landmarks are observed in a virtual "current" keyframe (index K) with their own descriptor
    K = len(targets)
    model.kf_descs.append(lms["desc"].copy())
    model.kf_lms.append(np.arange(m, dtype=np.int64))
    for i in range(m):
        model.obs[i][K] = i
    for lm in range(m, m + extra):
        model.compute_descriptor(lm)
    return model


@pytest.mark.parametrize("seed", range(3))
def test_adapter_protocol_equals_sequential_reference(plp, orc, seed):
    """mapping_module.cc:711-714: batched search + ordered effects + re-search == the landmark-by-landmark loop."""
    rng = np.random.default_rng(seed + 100)
    lms, targets = fuse_data.make_point_fuse_scene(seed + 20, m=300, num_targets=4, n_extra=80)
    for t in targets:
        t.pop("skip")   # is_observed_in_keyframe comes from the model below
    grid = plp.capi.make_grid(synth.COLS, synth.ROWS)
    cam = _cam(plp)
    sf, isg = synth.scale_factors(), fuse_data.inv_level_sigma_sq()

    def search_batch(kfs, js, descs):
        sub = {k: (np.asarray(v)[js] if k != "desc" else np.asarray(descs)) for k, v in lms.items()}
        return np.stack([orc.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, targets[k], sub, 3.0, 1)[0]
                         for k in kfs])

    def search_one(kf, js, descs):
        return search_batch([kf], js, descs)[0]

    lm_ids = [i for i in range(300) if lms["valid"][i]]
    js_of = {lm: lm for lm in lm_ids}

    def sb(kfs, js, descs):
        return search_batch(kfs, [lm_ids[j] for j in js], descs)

    def so(kf, js, descs):
        return search_one(kf, [lm_ids[j] for j in js], descs)

    a = _model_for(lms, targets, np.random.default_rng(seed))
    b = _model_for(lms, targets, np.random.default_rng(seed))
    order = list(range(len(targets)))
    na = fuse_data.fuse_sequential(a, lm_ids, order, so)
    nb, researches = fuse_data.fuse_batched(b, lm_ids, order, sb)
    assert na == nb and na > 50
    sa, sb_ = a.state(), b.state()
    assert all(np.array_equal(x, y) for x, y in zip(sa[0], sb_[0]))
    assert sa[1] == sb_[1]
    assert np.array_equal(sa[2], sb_[2]) and np.array_equal(sa[3], sb_[3])
    assert a.erased.sum() > 5          # replace() fired
    assert js_of and researches >= 0


@pytest.mark.parametrize("seed", range(3))
def test_fuse_points_oracle_matches_python_restatement(plp, orc, seed):
    """The C++ oracle against a second, independent restatement of fuse.cc (pure Python, libm logf via ctypes)."""
    stereo = bool(seed & 1)
    lms, targets = fuse_data.make_point_fuse_scene(seed + 50, m=260, num_targets=2, n_extra=60, stereo=stereo)
    grid = plp.capi.make_grid(synth.COLS, synth.ROWS)
    cam = _cam(plp, stereo)
    sf, isg = synth.scale_factors(), fuse_data.inv_level_sigma_sq()
    matched = 0
    for tgt in targets:
        for margin, mode in [(3.0, 1), (4.0, 0), (8.0, 1)]:
            o_idx, o_dist, _ = orc.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, tgt, lms, margin, mode)
            p_idx, p_dist = fuse_data.fuse_search_points_python(grid, cam, sf, isg, fuse_data.LOG_SF, tgt, lms, margin, mode)
            assert np.array_equal(o_idx, p_idx)
            assert np.array_equal(o_dist, p_dist)
            matched += (o_idx >= 0).sum()
    assert matched > 60


@pytest.mark.parametrize("seed", range(2))
def test_fuse_lines_oracle_matches_python_restatement(plp, orc, seed):
    levels = 1 + 2 * (seed & 1)
    lms, targets = fuse_data.make_line_fuse_scene(seed + 30, m=160, num_targets=2, n_extra=30, num_levels=levels)
    cam = _cam(plp)
    sf = np.float32([1.0, 2.0, 4.0])[:levels]
    isg = (1.0 / (sf * sf)).astype(np.float32)
    lsf = float(np.log(np.float32(2.0)).astype(np.float32))
    matched = 0
    for tgt in targets:
        for margin in (10.0, 4.0):
            o_idx, o_dist, _ = orc.fuse_search_lines(cam, sf, isg, lsf, tgt, lms, margin)
            p_idx, p_dist = fuse_data.fuse_search_lines_python(cam, sf, isg, lsf, tgt, lms, margin)
            assert np.array_equal(o_idx, p_idx)
            assert np.array_equal(o_dist, p_dist)
            matched += (o_idx >= 0).sum()
    assert matched > 20
