"""CPU-oracle backend of plpslam_b200.sequence.SequentialTracker (TEST INFRASTRUCTURE: the product never imports it)."""
import numpy as np

import oracle_api
import synth


class OracleBackend:
    def __init__(self, orc=None):
        self.orc = orc or oracle_api.Oracle()
        self.p = oracle_api.orb_params()
        self.scale_factors = synth.scale_factors()
        self.inv_level_sigma_sq = synth.inv_level_sigma_sq()

    def extract(self, img):
        r = self.orc.orb_extract(self.p, img)
        return r["kps"], r["desc"]

    def match_current_and_last_frames(self, grid, cam, curr, Tc, Tl, last, margin):
        return self.orc.match_current_and_last_frames(grid, self.scale_factors, cam, curr, Tc, Tl, last, margin, True)

    def match_frame_and_landmarks(self, grid, frm, q, margin, lowe_ratio):
        return self.orc.match_frame_and_landmarks(grid, self.scale_factors, frm, q, margin, lowe_ratio)

    def pose_optimize(self, cam, T, pts):
        T_out, pt_out, _, n_inl, _ = self.orc.pose_optimize(cam, T, pts)
        return T_out, pt_out, n_inl
