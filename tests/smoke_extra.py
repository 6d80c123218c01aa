"""Extra smoke checks appended as the hot path widens (called by __graft_entry__.smoke())."""
import numpy as np


def run(pkg, ctx, orc):
    import oracle_api
    import synth

    # ORB extraction of one small textured frame, bit-exact vs the oracle
    img = synth.make_texture(4321, 240, 320, n_rect=120, n_blob=500)
    ext = pkg.OrbExtractor(ctx, 240, 320, 500)
    kps, desc = ext.extract(img)
    r = orc.orb_extract(oracle_api.orb_params(500), img)
    assert len(kps) == len(r["kps"]) > 100, "ORB keypoint count differs from the oracle"
    for f in ("x", "y", "angle", "response", "octave"):
        assert np.array_equal(kps[f], r["kps"][f]), f"ORB field {f} differs from the oracle"
    assert np.array_equal(desc, r["desc"]), "ORB descriptors differ from the oracle"
    ext.close()

    # pose optimiser (points + lines), 1e-4 relative vs the oracle
    cam = pkg.capi.make_camera(synth.FX, synth.FY, synth.CX, synth.CY, synth.COLS, synth.ROWS)
    T_gt, T_init, pts, lines = synth.make_pose_opt_scene(5, n_pts=400, n_lines=60)
    To, po, lo, no, _ = orc.pose_optimize(cam, T_init, pts, lines)
    Tg, pg, lg, ng = ctx.pose_optimize(cam, T_init, pts, lines)
    assert np.linalg.norm(Tg - To) / np.linalg.norm(To) < 1e-4 and ng == no and np.array_equal(pg, po)
    # LSD + LBD line extraction of one small frame, bit-exact vs the oracle
    limg = synth.make_line_image(7, 240, 320, n_patch=16)
    trk = pkg.LineFeatureTracker(ctx, 240, 320)
    kl, lbd, fn = trk.extract_LSD_LBD(limg)
    okl, olbd, ofn = orc.line_extract(limg)
    assert len(kl) == len(okl) > 3 and np.array_equal(lbd, olbd) and np.array_equal(fn, ofn), "line extraction differs"
    for f in ("start_x", "start_y", "end_x", "end_y", "angle", "num_pixels"):
        assert np.array_equal(kl[f], okl[f]), f"KeyLine field {f} differs from the oracle"
    trk.close()
    print(f"smoke extras ok: {len(kl)} LSD/LBD lines bit-exact")
    print(f"smoke extras ok: {len(kps)} ORB keypoints bit-exact, pose-opt inliers {ng}")

    # mapping-thread matchers: fuse search, DBoW2 transform, essential RANSAC -- tiny cases, bit-exact vs the oracle
    import bow_data
    import ess_data
    import fuse_data
    grid = pkg.capi.make_grid(synth.COLS, synth.ROWS)
    lms, targets = fuse_data.make_point_fuse_scene(3, m=200, num_targets=2, n_extra=50)
    sf, isg = synth.scale_factors(), fuse_data.inv_level_sigma_sq()
    g_idx, _ = ctx.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, targets, lms, 3.0, 1)
    for t, tgt in enumerate(targets):
        assert np.array_equal(g_idx[t], orc.fuse_search_points(grid, cam, sf, isg, fuse_data.LOG_SF, tgt, lms, 3.0, 1)[0]), \
            "fuse search differs from the oracle"
    vocab = bow_data.make_vocab(7, k=10, L=3)
    rng = np.random.default_rng(1)
    d = synth.rand_desc(rng, 300)
    ov = orc.bow_vocab_create(10, 3, vocab["parent"], vocab["desc"], vocab["weight"], vocab["is_leaf"])
    gv = pkg.BowVocabulary(ctx, k=10, L=3, parent=vocab["parent"], desc=vocab["desc"], weight=vocab["weight"],
                           is_leaf=vocab["is_leaf"])
    assert all(np.array_equal(a, b) for a, b in zip(gv.transform(d, 1), orc.bow_transform(ov, d, 1))), "BoW transform differs"
    gv.close()
    orc.bow_vocab_destroy(ov)
    b1, b2, matches, _ = ess_data.make_two_view(11, n=120)
    smp = ess_data.draw_samples(2, len(matches), 20)
    want, got = orc.essential_ransac(b1, b2, matches, smp, False), ctx.essential_ransac(b1, b2, matches, smp, False)
    assert got[0] == want[0] and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]), "essential RANSAC differs"
    print(f"smoke extras ok: fuse {int((g_idx >= 0).sum())} matches, BoW transform, essential RANSAC ({int(got[1].sum())} inliers) bit-exact")
