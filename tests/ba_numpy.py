"""Independent numpy restatement of the g2o machinery behind the reference's optimisers (TEST INFRASTRUCTURE).

g2o is an un-vendored, unversioned third-party dependency of the reference and absent from this environment, so the C++
oracle (oracle/g2o_lite.hpp, oracle/local_ba.cc, oracle/pose_opt.cc) cannot be pinned against g2o itself.  This module is
a SECOND restatement, written separately and in a different formulation, that the C++ oracle has to match try for try:

  * poses are 4x4 matrices updated with T <- Exp(delta) T (optimize/g2o/se3/shot_vertex.h:58-62, SE3Quat::exp written with
    rotation matrices, no quaternions);
  * the linear system is the FULL (poses + landmarks) normal-equation matrix, assembled densely and solved in one piece --
    no Schur complement, no per-landmark block inverses (g2o's BlockSolver eliminates the marginalised landmark blocks;
    the result of an exact elimination equals the full solve);
  * every Jacobian is numeric (central differences through the vertex oplus) except where stated: the analytic point
    Jacobians of the reference (optimize/g2o/se3/perspective_pose_opt_edge.cc:76-101, perspective_reproj_edge.cc:78-125)
    are checked against these differences in tests/test_ba_oracle.py;
  * OptimizationAlgorithmLevenberg (computeLambdaInit tau = 1e-5, rho = (chi_old - chi_new) / (dx.(lambda dx + b) + 1e-3),
    scale factor clamp [1/3, 2/3], nu doubling, <= 10 trials after failure, termination rules), RobustKernelHuber and the
    level / outlier bookkeeping of optimize/local_bundle_adjuster.cc:276-372 are restated from SURVEY.md Appendix B.
"""
from __future__ import annotations

import numpy as np

CHI2_2D = float(np.float32(5.99146))
CHI2_3D = float(np.float32(7.81473))


# ------------------------------------------------------------------------------------------------------------------
# SE3 (g2o::SE3Quat::exp, shot_vertex::oplusImpl)
# ------------------------------------------------------------------------------------------------------------------
def skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def se3_exp(u):
    omega, ups = np.asarray(u[:3], float), np.asarray(u[3:], float)
    theta = np.linalg.norm(omega)
    O = skew(omega)
    O2 = O @ O
    if theta < 1e-5:
        a, b, c = 1.0, 0.5, 1.0 / 6.0
    else:
        a = np.sin(theta) / theta
        b = (1 - np.cos(theta)) / theta ** 2
        c = (theta - np.sin(theta)) / theta ** 3
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + a * O + b * O2
    T[:3, 3] = (np.eye(3) + b * O + c * O2) @ ups
    return T


def pose_oplus(T, u):
    return se3_exp(u) @ T


# ------------------------------------------------------------------------------------------------------------------
# Line3D (optimize/g2o/line3d.h:116-186): Pluecker (w, d) <-> orthonormal (U, W), 4-dof update
# ------------------------------------------------------------------------------------------------------------------
def quat_to_R(w, x, y, z):
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def line_oplus(L, v):
    w, d = np.asarray(L[:3], float), np.asarray(L[3:], float)
    nd, nw = np.linalg.norm(d), np.linalg.norm(w)
    wn = 1.0 / np.hypot(nd, nw)
    W = np.array([[nw * wn, -nd * wn], [nd * wn, nw * wn]])
    c = np.cross(w, d)
    U = np.stack([w / nw, d / nd, c / np.linalg.norm(c)], 1)
    q = np.array([np.sqrt(1 - (v[0] ** 2 + v[1] ** 2 + v[2] ** 2)), v[0], v[1], v[2]])
    q = q / np.linalg.norm(q)
    U = U @ quat_to_R(*q)
    W = W @ np.array([[np.cos(v[3]), -np.sin(v[3])], [np.sin(v[3]), np.cos(v[3])]])
    out = np.concatenate([U[:, 0] * W[0, 0], U[:, 1] * W[1, 0]])
    out = out / np.linalg.norm(out[3:])
    return out / np.linalg.norm(out[3:])


# ------------------------------------------------------------------------------------------------------------------
# edge residuals
# ------------------------------------------------------------------------------------------------------------------
def point_residual(cam, T, X, obs):
    """obs = (x, y, x_right); x_right < 0 -> monocular 2-vector (perspective_pose_opt_edge.h:55-113)."""
    fx, fy, cx, cy, bf = cam
    pc = T[:3, :3] @ X + T[:3, 3]
    rx = fx * pc[0] / pc[2] + cx
    e = [obs[0] - rx, obs[1] - (fy * pc[1] / pc[2] + cy)]
    if not (obs[2] < 0):
        e.append(obs[2] - (rx - bf / pc[2]))
    return np.array(e)


def line_residual(cam, T, L, obs):
    """pose_opt_edge_line3d_orthonormal.h:61-89 / reproj_edge_line3d_orthonormal.h:62-90."""
    fx, fy, cx, cy, _ = cam
    R, t = T[:3, :3], T[:3, 3]
    lc = R @ L[:3] + skew(t) @ R @ L[3:]
    K = np.array([[fy, 0, 0], [0, fx, 0], [-fy * cx, -fx * cy, fx * fy]])
    p = K @ lc
    den = np.hypot(p[0], p[1])
    return np.array([(obs[0] * p[0] + obs[1] * p[1] + p[2]) / den, (obs[2] * p[0] + obs[3] * p[1] + p[2]) / den])


def plane_residual(X, fn):
    return (X @ fn[:3] + fn[3]) / np.linalg.norm(fn[:3])


def huber(e2, delta):
    d2 = delta * delta
    if e2 <= d2:
        return e2, 1.0
    s = np.sqrt(e2)
    return 2 * s * delta - d2, delta / s


def numeric_jacobian(f, dim, delta=1e-9):
    cols = []
    for d in range(dim):
        u = np.zeros(dim)
        u[d] = delta
        ep = f(u)
        u[d] = -delta
        em = f(u)
        cols.append((ep - em) / (2 * delta))
    return np.stack(cols, 1)


def point_jac_pose_analytic(cam, T, X, stereo):
    """The reference's formulas (perspective_pose_opt_edge.cc:76-101, :142-173) transcribed."""
    fx, fy, cx, cy, bf = cam
    x, y, z = T[:3, :3] @ X + T[:3, 3]
    z2 = z * z
    J = np.array([[x * y / z2 * fx, -(1 + x * x / z2) * fx, y / z * fx, -1 / z * fx, 0, x / z2 * fx],
                  [(1 + y * y / z2) * fy, -x * y / z2 * fy, -x / z * fy, 0, -1 / z * fy, y / z2 * fy]])
    if stereo:
        J = np.vstack([J, [J[0, 0] - bf * y / z2, J[0, 1] + bf * x / z2, J[0, 2], J[0, 3], 0, J[0, 5] - bf / z2]])
    return J


def point_jac_landmark_analytic(cam, T, X, stereo):
    """perspective_reproj_edge.cc:78-125 (:166-214 stereo)."""
    fx, fy, cx, cy, bf = cam
    R = T[:3, :3]
    x, y, z = R @ X + T[:3, 3]
    z2 = z * z
    J = np.stack([-fx * R[0] / z + fx * x * R[2] / z2, -fy * R[1] / z + fy * y * R[2] / z2])
    if stereo:
        J = np.vstack([J, J[0] - bf * R[2] / z2])
    return J


# ------------------------------------------------------------------------------------------------------------------
# the graph + Levenberg-Marquardt on the full system
# ------------------------------------------------------------------------------------------------------------------
class Graph:
    def __init__(self, prob, cam, numeric_point_jacobians=False):
        """prob: tests/ba_data.BAProblem"""
        self.cam = cam
        self.stereo_setup = bool(prob.stereo)
        self.T = [np.array(t, float).reshape(4, 4) for t in prob.kf_pose_cw.reshape(-1, 4, 4)]
        self.fixed = np.asarray(prob.kf_fixed).astype(bool)
        self.X = np.array(prob.pt_pos_w, float)
        self.L = np.array(prob.line_plucker, float)
        self.pe = [dict(kf=int(k), lm=int(l), obs=np.array(o, float), info=float(i), level=0, robust=True, e=None)
                   for k, l, o, i in zip(prob.pt_edge_kf, prob.pt_edge_lm, prob.pt_edge_obs, prob.pt_edge_inv_sigma_sq)]
        self.le = [dict(kf=int(k), lm=int(l), obs=np.array(o, float), info=float(i), level=0, robust=True, e=None)
                   for k, l, o, i in zip(prob.line_edge_kf, prob.line_edge_lm, prob.line_edge_obs, prob.line_edge_inv_sigma_sq)]
        self.pl = [dict(lm=int(l), fn=np.array(f, float), e=None) for l, f in zip(prob.plane_edge_lm, prob.plane_edge_fn)]
        self.delta_pt = float(np.sqrt(np.float32(5.99146))) if not prob.stereo else float(np.sqrt(np.float32(7.81473)))
        self.delta_ln = float(np.sqrt(np.float32(5.99146)))
        self.numeric_pt = numeric_point_jacobians
        self.lm_tries = 0
        self.trace = []   # (lambda, rho, accepted) per try
        # unknown layout: free poses (6 each), then active points (3), then active lines (4) -- rebuilt per optimize()
        self.hidx = np.full(len(self.T), -1)
        self.hidx[~self.fixed] = np.arange((~self.fixed).sum())

    # ---- errors
    def compute_errors(self):
        for e in self.pe:
            if e["level"] == 0:
                e["e"] = point_residual(self.cam, self.T[e["kf"]], self.X[e["lm"]], e["obs"])
        for e in self.le:
            if e["level"] == 0:
                e["e"] = line_residual(self.cam, self.T[e["kf"]], self.L[e["lm"]], e["obs"])
        for e in self.pl:
            e["e"] = plane_residual(self.X[e["lm"]], e["fn"])

    def robust_chi2(self):
        chi = 0.0
        for e in self.pe:
            if e["level"] == 0:
                c2 = e["info"] * float(e["e"] @ e["e"])
                chi += huber(c2, self.delta_pt)[0] if e["robust"] else c2
        for e in self.le:
            if e["level"] == 0:
                c2 = e["info"] * float(e["e"] @ e["e"])
                chi += huber(c2, self.delta_ln)[0] if e["robust"] else c2
        for e in self.pl:
            chi += huber(e["e"] ** 2, 1.0)[0]
        return chi

    # ---- full normal equations
    def build(self):
        nf = int((~self.fixed).sum())
        act_p = np.zeros(len(self.X), bool)
        act_l = np.zeros(len(self.L), bool)
        for e in self.pe:
            if e["level"] == 0:
                act_p[e["lm"]] = True
        for e in self.pl:
            act_p[e["lm"]] = True
        for e in self.le:
            if e["level"] == 0:
                act_l[e["lm"]] = True
        off_p = np.full(len(self.X), -1)
        off_p[act_p] = 6 * nf + 3 * np.arange(act_p.sum())
        off_l = np.full(len(self.L), -1)
        off_l[act_l] = 6 * nf + 3 * act_p.sum() + 4 * np.arange(act_l.sum())
        n = 6 * nf + 3 * int(act_p.sum()) + 4 * int(act_l.sum())
        H = np.zeros((n, n))
        b = np.zeros(n)

        def add(blocks, w, r):
            """blocks: list of (offset, J); H += J^T w J, b += -J^T w r."""
            for oa, Ja in blocks:
                b[oa:oa + Ja.shape[1]] += Ja.T @ (-w * r)
                for ob, Jb in blocks:
                    H[oa:oa + Ja.shape[1], ob:ob + Jb.shape[1]] += Ja.T @ (w * Jb)

        for e in self.pe:
            if e["level"]:
                continue
            T, X, stereo = self.T[e["kf"]], self.X[e["lm"]], not (e["obs"][2] < 0)
            if self.numeric_pt:
                Jp = numeric_jacobian(lambda u: point_residual(self.cam, pose_oplus(T, u), X, e["obs"]), 6)
                Jl = numeric_jacobian(lambda u: point_residual(self.cam, T, X + u, e["obs"]), 3)
            else:
                Jp = point_jac_pose_analytic(self.cam, T, X, stereo)
                Jl = point_jac_landmark_analytic(self.cam, T, X, stereo)
            w = e["info"]
            if e["robust"]:
                w *= huber(e["info"] * float(e["e"] @ e["e"]), self.delta_pt)[1]
            blocks = [(off_p[e["lm"]], Jl)]
            if self.hidx[e["kf"]] >= 0:
                blocks.append((6 * self.hidx[e["kf"]], Jp))
            add(blocks, w, e["e"])
        for e in self.le:
            if e["level"]:
                continue
            T, L = self.T[e["kf"]], self.L[e["lm"]]
            Jp = numeric_jacobian(lambda u: line_residual(self.cam, pose_oplus(T, u), L, e["obs"]), 6)
            Jl = numeric_jacobian(lambda u: line_residual(self.cam, T, line_oplus(L, u), e["obs"]), 4)
            w = e["info"]
            if e["robust"]:
                w *= huber(e["info"] * float(e["e"] @ e["e"]), self.delta_ln)[1]
            blocks = [(off_l[e["lm"]], Jl)]
            if self.hidx[e["kf"]] >= 0:
                blocks.append((6 * self.hidx[e["kf"]], Jp))
            add(blocks, w, e["e"])
        for e in self.pl:
            X = self.X[e["lm"]]
            J = numeric_jacobian(lambda u: np.array([plane_residual(X + u, e["fn"])]), 3)
            w = huber(e["e"] ** 2, 1.0)[1]
            add([(off_p[e["lm"]], J)], w, np.array([e["e"]]))
        return H, b, off_p, off_l, nf

    def apply(self, x, off_p, off_l, nf):
        for k in range(len(self.T)):
            if self.hidx[k] >= 0:
                self.T[k] = pose_oplus(self.T[k], x[6 * self.hidx[k]:6 * self.hidx[k] + 6])
        for l in np.nonzero(off_p >= 0)[0]:
            self.X[l] = self.X[l] + x[off_p[l]:off_p[l] + 3]
        for l in np.nonzero(off_l >= 0)[0]:
            self.L[l] = line_oplus(self.L[l], x[off_l[l]:off_l[l] + 4])

    def optimize(self, iterations):
        lam, ni = 0.0, 2.0
        done = 0
        for it in range(iterations):
            self.compute_errors()
            current = self.robust_chi2()
            H, b, off_p, off_l, nf = self.build()
            if it == 0:
                lam, ni = 1e-5 * np.abs(np.diag(H)).max(), 2.0
            rho, qmax, finite = 0.0, 0, True
            while True:
                backup = ([t.copy() for t in self.T], self.X.copy(), self.L.copy())
                ok = True
                try:
                    A = H + lam * np.eye(len(b))
                    np.linalg.cholesky(A)   # the block solver fails on a non positive definite system
                    x = np.linalg.solve(A, b)
                except np.linalg.LinAlgError:
                    ok, x = False, np.zeros(len(b))
                if ok:
                    self.apply(x, off_p, off_l, nf)
                self.compute_errors()
                temp = self.robust_chi2() if ok else np.finfo(float).max
                scale = float(x @ (lam * x + b)) + 1e-3
                rho = (current - temp) / scale
                self.lm_tries += 1
                accepted = rho > 0 and np.isfinite(temp)
                self.trace.append((lam, rho, accepted))
                if accepted:
                    alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
                    lam *= max(1.0 / 3.0, alpha)
                    ni = 2.0
                    current = temp
                else:
                    lam *= ni
                    ni *= 2
                    self.T, self.X, self.L = backup
                    if not np.isfinite(lam):
                        finite = False
                        break
                qmax += 1
                if not (rho < 0 and qmax < 10):
                    break
            done += 1
            if qmax == 10 or rho == 0 or not finite:
                break
        return done

    # ---- optimize/local_bundle_adjuster.cc:284-372 (+ line variant)
    def _pt_outlier(self, e):
        pc = self.T[e["kf"]][:3, :3] @ self.X[e["lm"]] + self.T[e["kf"]][:3, 3]
        thr = CHI2_2D if e["obs"][2] < 0 else CHI2_3D
        return thr < e["info"] * float(e["e"] @ e["e"]) or not (0.0 < pc[2])

    def local_ba(self, num_first=5, num_second=10, line_depth_positive=None):
        it1 = self.optimize(num_first)
        for e in self.pe:
            if self._pt_outlier(e):
                e["level"] = 1
            e["robust"] = False
        for e in self.le:
            bad_depth = line_depth_positive is not None and not line_depth_positive(self, e)
            if CHI2_2D < e["info"] * float(e["e"] @ e["e"]) or bad_depth:
                e["level"] = 1
            e["robust"] = False
        it2 = self.optimize(num_second)
        pt_out = np.array([self._pt_outlier(e) for e in self.pe], np.uint8)
        ln_out = np.array([CHI2_2D < e["info"] * float(e["e"] @ e["e"]) or
                           (line_depth_positive is not None and not line_depth_positive(self, e)) for e in self.le], np.uint8)
        return it1, it2, pt_out, ln_out
