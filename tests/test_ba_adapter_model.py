"""The local-BA adapter protocol (gather -> flatten -> solve -> outlier erase + write-back + end-point trimming) on the map
model of tests/ba_adapter_model.py: with the oracle as the solver on the CPU, with plp_local_ba on the GPU, and both must
leave the map in the same state."""
import ctypes as C

import numpy as np
import pytest

import ba_adapter_model as bam
import ba_data
import synth

_P = C.c_void_p


def _oracle_solve(orc):
    def solve(prob):
        r = ba_data.oracle_local_ba(orc, prob)
        return dict(kf_pose_cw=r.kf_pose_cw.reshape(-1, 4, 4), pt_pos_w=r.pt_pos_w, line_plucker=r.line_plucker,
                    pt_edge_outlier=r.pt_edge_outlier, line_edge_outlier=r.line_edge_outlier, lm_tries=r.lm_tries)
    return solve


def _oracle_trim(orc):
    orc.lib.orc_endpoint_trimming.restype = C.c_int

    def trim(cam4, pose, plucker, sp, ep, old, md):
        out = np.zeros(6)
        d = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(_P)
        f = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(_P)
        keep = orc.lib.orc_endpoint_trimming(d(cam4), d(pose), d(plucker), f(sp), f(ep), d(old), C.c_double(md),
                                             out.ctypes.data_as(_P))
        return bool(keep), out
    return trim


def _make_map(seed=11, n_local=6, n_fixed=3):
    prob = ba_data.make_ba_problem(seed, n_local=n_local, n_fixed=n_fixed, n_points=250, n_lines=50, n_plane_pts=0)
    ep = prob.gt["line_endpoints"]   # Line::get_pos_in_world(): the 3-D end points the Pluecker coordinates came from
    return prob, bam.map_from_problem(prob, n_local, endpoints=ep)


def test_gather_reproduces_the_window_and_write_back_updates_the_map(orc):
    prob, (kfs, lms, lines) = _make_map()
    poses_before = [k.pose.copy() for k in kfs]
    n_obs_before = sum(len(l.obs) for l in lms) + sum(len(l.obs) for l in lines)
    out = bam.local_bundle_adjust(kfs[0], _oracle_solve(orc), _oracle_trim(orc))
    # [1] local = current + covisibilities, fixed = the other observers; the flattened problem is the generator's
    assert out["n_local"] == 6 and out["n_fixed"] == 3
    g = out["problem"]
    assert np.array_equal(g.kf_fixed, prob.kf_fixed) and np.array_equal(g.kf_pose_cw, prob.kf_pose_cw)
    assert np.array_equal(g.pt_edge_lm, prob.pt_edge_lm) and np.array_equal(g.pt_edge_kf, prob.pt_edge_kf)
    assert np.array_equal(g.pt_edge_obs, prob.pt_edge_obs) and np.array_equal(g.line_edge_obs, prob.line_edge_obs)
    assert np.all(np.diff(g.pt_edge_lm) >= 0) and np.all(np.diff(g.line_edge_lm) >= 0)   # grouped by landmark
    # the solve equals the oracle on the generator's arrays
    r = ba_data.oracle_local_ba(orc, prob)
    assert np.array_equal(out["result"]["kf_pose_cw"].reshape(-1, 16), r.kf_pose_cw.reshape(-1, 16))
    # [8] fixed keyframes keep their pose, local ones moved, landmark positions are the solver's
    for k, kf in enumerate(kfs):
        moved = not np.array_equal(kf.pose, poses_before[k])
        assert moved == (prob.kf_fixed[k] == 0)
    assert all(np.array_equal(lm.pos, r.pt_pos_w[i]) for i, lm in enumerate(lms))
    # [7] every outlier observation is gone from BOTH tables, every inlier one is still there
    n_out = int(r.pt_edge_outlier.sum() + r.line_edge_outlier.sum())
    assert out["n_erased_obs"] == n_out and n_out > 0
    assert sum(len(l.obs) for l in lms) + sum(len(l.obs) for l in lines) == n_obs_before - n_out
    for lm in lms:
        for kf, slot in lm.obs.items():
            assert kf.landmarks[slot] is lm
    for kf in kfs:
        for slot, lm in enumerate(kf.landmarks):
            assert lm is None or lm.obs.get(kf) == slot
    # lines: Pluecker written; trimmed end points lie on the optimised line; the rest is prepared for erasing
    kept = [l for l in lines if not l.erased]
    assert len(kept) >= 0.6 * len(lines) and out["n_lines_erased"] == len(lines) - len(kept)
    for i, ll in enumerate(lines):
        assert np.array_equal(ll.plucker, r.line_plucker[i])
    for ll in kept:
        m, dd = ll.plucker[:3], ll.plucker[3:]
        for X in (ll.endpoints[:3], ll.endpoints[3:]):     # X on the line  <=>  X x d = m
            assert np.linalg.norm(np.cross(X, dd) - m) < 1e-6 * (1 + np.linalg.norm(m))


def test_keyframe_id_zero_is_fixed_even_inside_the_window_and_erased_keyframes_are_skipped(orc):
    prob, (kfs, lms, lines) = _make_map(seed=12)
    kfs[1].id = 0                      # local_bundle_adjuster.cc:197-202
    kfs[2].erased = True               # will_be_erased(): not a local keyframe, and its observations produce no edge
    out = bam.local_bundle_adjust(kfs[0], _oracle_solve(orc), _oracle_trim(orc))
    g = out["problem"]
    ks = out["keyframes"]
    assert kfs[2] not in ks and len(ks) == len(kfs) - 1
    # a fixed local keyframe is written back too (set_cam_pose(vertex estimate), :382-388): its pose after the
    # matrix -> quaternion -> matrix round trip of the vertex
    assert g.kf_fixed[ks.index(kfs[1])] == 1 and np.allclose(kfs[1].pose, prob.kf_pose_cw[1].reshape(4, 4), atol=1e-12)
    # a landmark stays local only if a surviving LOCAL keyframe observes it; its edge on the erased keyframe is dropped
    local_ids = [k for k in range(6) if k != 2]
    is_local_lm = np.zeros(len(prob.pt_pos_w), bool)
    is_local_lm[prob.pt_edge_lm[np.isin(prob.pt_edge_kf, local_ids)]] = True
    assert len(g.pt_edge_kf) == int((is_local_lm[prob.pt_edge_lm] & (prob.pt_edge_kf != 2)).sum())


@pytest.mark.gpu
def test_gpu_solver_leaves_the_map_in_the_oracle_state(ctx, orc, plp):
    from plpslam_b200.ba import LocalBA

    def gpu_solve(prob):
        st = prob.struct()
        ba = LocalBA(ctx, st, (len(prob.kf_fixed), len(prob.pt_pos_w), len(prob.line_plucker), len(prob.pt_edge_kf),
                               len(prob.line_edge_kf)))
        out = ba.solve()
        ba.close()
        return out

    _, (kfs_o, lms_o, lines_o) = _make_map()
    _, (kfs_g, lms_g, lines_g) = _make_map()
    o = bam.local_bundle_adjust(kfs_o[0], _oracle_solve(orc), _oracle_trim(orc))
    g = bam.local_bundle_adjust(kfs_g[0], gpu_solve, _oracle_trim(orc))
    assert g["n_erased_obs"] == o["n_erased_obs"] and g["result"]["lm_tries"] == o["result"]["lm_tries"]
    for a, b in zip(kfs_g, kfs_o):
        assert np.linalg.norm(a.pose - b.pose) / np.linalg.norm(b.pose) < 1e-4
        assert [l is None for l in a.landmarks] == [l is None for l in b.landmarks]
    for a, b in zip(lms_g, lms_o):
        assert np.linalg.norm(a.pos - b.pos) / np.linalg.norm(b.pos) < 1e-4
    assert sum(a.erased != b.erased for a, b in zip(lines_g, lines_o)) <= 1
