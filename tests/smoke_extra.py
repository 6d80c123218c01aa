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
