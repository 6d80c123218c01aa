"""Map-state model of the local-BA adapter (structure-plp-slam_b200/host/plpslam_b200_adapter.hpp::local_bundle_adjust),
the replacement of optimize::local_bundle_adjuster[_extended_line]::optimize (optimize/local_bundle_adjuster.cc:62-410,
local_bundle_adjuster_extended_line.cc:69-674).  The C++ adapter can only be syntax-checked here (the reference's
OpenCV / Eigen / g2o dependencies are absent), so its PROTOCOL is modelled on plain Python objects that carry the
reference's tables -- keyframe -> landmark slots, landmark -> {keyframe: keypoint index} observations, the covisibility
list -- and executed with the oracle (CPU) or the C ABI (GPU) as the solver:

    gather      local keyframes = current + covisibilities, local landmarks = everything they observe, fixed keyframes =
                other observers of those landmarks (local_bundle_adjuster.cc:72-158); ascending ids
    flatten     keyframes (local first, id 0 fixed), one edge per observation grouped by landmark
    solve       plp_local_ba / orc_local_ba
    write back  outlier observations erased on both sides, poses, positions, Pluecker coordinates, end-point trimming on
                the reference keyframe, lines that fail it prepared for erasing (:375-409, :642-672)
"""
from __future__ import annotations

import numpy as np

import ba_data
import synth


class Keyframe:
    def __init__(self, kid, pose, n_slots=0):
        self.id = kid
        self.pose = np.array(pose, float).reshape(4, 4)
        self.kp = []          # (x, y, x_right, inv_sigma_sq) per keypoint slot
        self.landmarks = []   # landmark or None per keypoint slot
        self.kl = []          # (sx, sy, ex, ey, inv_sigma_sq) per keyline slot
        self.lines = []
        self.covis = []
        self.erased = False


class Landmark:
    def __init__(self, lid, pos):
        self.id, self.pos, self.obs, self.erased = lid, np.array(pos, float), {}, False   # obs: keyframe -> slot


class Line:
    def __init__(self, lid, plucker, endpoints):
        self.id, self.plucker, self.endpoints = lid, np.array(plucker, float), np.array(endpoints, float)
        self.obs, self.erased, self.ref_kf = {}, False, None


def map_from_problem(prob: ba_data.BAProblem, n_local: int, endpoints=None):
    """A map whose local window around keyframe 0 is exactly `prob` (keyframes [0, n_local) covisible with keyframe 0)."""
    kfs = [Keyframe(k + 1, prob.kf_pose_cw[k]) for k in range(len(prob.kf_fixed))]   # ids 1..: no keyframe id 0
    lms = [Landmark(100 + i, p) for i, p in enumerate(prob.pt_pos_w)]
    for e in range(len(prob.pt_edge_kf)):
        kf, lm = kfs[prob.pt_edge_kf[e]], lms[prob.pt_edge_lm[e]]
        kf.kp.append((*prob.pt_edge_obs[e], prob.pt_edge_inv_sigma_sq[e]))
        kf.landmarks.append(lm)
        lm.obs[kf] = len(kf.kp) - 1
    lines = []
    for i, L in enumerate(prob.line_plucker):
        ep = endpoints[i] if endpoints is not None else np.zeros(6)
        lines.append(Line(500 + i, L, ep))
    for e in range(len(prob.line_edge_kf)):
        kf, ll = kfs[prob.line_edge_kf[e]], lines[prob.line_edge_lm[e]]
        kf.kl.append((*prob.line_edge_obs[e], prob.line_edge_inv_sigma_sq[e]))
        kf.lines.append(ll)
        ll.obs[kf] = len(kf.kl) - 1
        if ll.ref_kf is None:
            ll.ref_kf = kf
    kfs[0].covis = kfs[1:n_local]
    return kfs, lms, lines


def local_bundle_adjust(curr_kf: Keyframe, solve, trim, with_lines=True, stereo=False):
    """The adapter, step by step.  solve(problem) -> result with the plp_ba_result fields; trim(cam4, pose, plucker, sp, ep,
    old_endpoints, median_depth) -> (keep, new_endpoints)."""
    # [1] gather
    local = {curr_kf.id: curr_kf}
    for kf in curr_kf.covis:
        if kf is not None and not kf.erased:
            local[kf.id] = kf
    local_lms, local_lines = {}, {}
    for kf in local.values():
        for lm in kf.landmarks:
            if lm is not None and not lm.erased:
                local_lms.setdefault(lm.id, lm)
        if with_lines:
            for ll in kf.lines:
                if ll is not None and not ll.erased:
                    local_lines.setdefault(ll.id, ll)
    fixed = {}
    for tab in (local_lms, local_lines):
        for lm in tab.values():
            for kf in lm.obs:
                if kf is not None and not kf.erased and kf.id not in local:
                    fixed.setdefault(kf.id, kf)
    # [3-4] flatten (ascending ids, local first)
    kfs = [local[k] for k in sorted(local)] + [fixed[k] for k in sorted(fixed)]
    kf_index = {kf: i for i, kf in enumerate(kfs)}
    kf_fixed = np.array([1 if (i >= len(local) or kf.id == 0) else 0 for i, kf in enumerate(kfs)], np.uint8)
    lms = [local_lms[k] for k in sorted(local_lms)]
    lines = [local_lines[k] for k in sorted(local_lines)]
    pe_kf, pe_lm, pe_obs, pe_info, pe_owner = [], [], [], [], []
    for li, lm in enumerate(lms):
        for kf, slot in lm.obs.items():
            if kf is None or kf.erased:
                continue
            x, y, xr, info = kf.kp[slot]
            pe_kf.append(kf_index[kf]); pe_lm.append(li); pe_obs.append((x, y, xr)); pe_info.append(info); pe_owner.append((kf, lm))
    le_kf, le_lm, le_obs, le_info, le_owner = [], [], [], [], []
    for li, ll in enumerate(lines):
        for kf, slot in ll.obs.items():
            if kf is None or kf.erased:
                continue
            sx, sy, ex, ey, info = kf.kl[slot]
            le_kf.append(kf_index[kf]); le_lm.append(li); le_obs.append((sx, sy, ex, ey)); le_info.append(info); le_owner.append((kf, ll))
    prob = ba_data.BAProblem(stereo=stereo, kf_pose_cw=np.stack([k.pose for k in kfs]), kf_fixed=kf_fixed,
                             pt_pos_w=np.stack([l.pos for l in lms]) if lms else np.zeros((0, 3)),
                             pt_edge_kf=pe_kf, pt_edge_lm=pe_lm, pt_edge_obs=np.array(pe_obs, np.float32).reshape(-1, 3),
                             pt_edge_inv_sigma_sq=pe_info,
                             line_plucker=np.stack([l.plucker for l in lines]) if lines else np.zeros((0, 6)),
                             line_edge_kf=le_kf, line_edge_lm=le_lm, line_edge_obs=np.array(le_obs, np.float32).reshape(-1, 4),
                             line_edge_inv_sigma_sq=le_info)
    res = solve(prob)
    # [7-8] write back
    n_erased = 0
    for e, (kf, lm) in enumerate(pe_owner):
        if res["pt_edge_outlier"][e] and not lm.erased:
            kf.landmarks[lm.obs[kf]] = None     # keyfrm->erase_landmark(lm)
            del lm.obs[kf]                      # lm->erase_observation(keyfrm)
            n_erased += 1
    for e, (kf, ll) in enumerate(le_owner):
        if res["line_edge_outlier"][e] and not ll.erased:
            kf.lines[ll.obs[kf]] = None
            del ll.obs[kf]
            n_erased += 1
    for kf in local.values():
        kf.pose = np.array(res["kf_pose_cw"][kf_index[kf]], float).reshape(4, 4)
    for li, lm in enumerate(lms):
        lm.pos = np.array(res["pt_pos_w"][li], float)
    cam4 = np.array([synth.FX, synth.FY, synth.CX, synth.CY])
    n_trim_erased = 0
    for li, ll in enumerate(lines):
        ll.plucker = np.array(res["line_plucker"][li], float)
        ref = ll.ref_kf
        slot = ll.obs.get(ref, -1)      # get_index_in_keyframe: -1 once the reference observation has been erased
        keep = slot != -1
        if keep:
            sx, sy, ex, ey, _ = ref.kl[slot]
            keep, new_ep = trim(cam4, ref.pose, ll.plucker, (sx, sy), (ex, ey), ll.endpoints, 6.0)
        if keep:
            ll.endpoints = new_ep
        else:
            ll.erased = True            # prepare_for_erasing()
            n_trim_erased += 1
    return dict(problem=prob, result=res, n_local=len(local), n_fixed=len(fixed), n_erased_obs=n_erased,
                n_lines_erased=n_trim_erased, keyframes=kfs, landmarks=lms, lines=lines)
