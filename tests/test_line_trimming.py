"""local_bundle_adjuster_extended_line::endpoint_trimming (optimize/local_bundle_adjuster_extended_line.cc:676-787): the
product's host function (structure-plp-slam_b200/host/plpslam_b200_line_trimming.h, compiled here with g++) against the C++
oracle and a numpy restatement with explicit matrices -- geometry first: a 3-D segment seen in a keyframe is recovered from
its Pluecker line and its own 2-D end points."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

import synth

ROOT = Path(__file__).resolve().parent.parent
_P = C.c_void_p


@pytest.fixture(scope="module")
def trim_lib(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    d = tmp_path_factory.mktemp("trim")
    src = d / "trim.cc"
    src.write_text('#include "plpslam_b200_line_trimming.h"\n'
                   'extern "C" int trim(const double *cam4, const double *pose, const double *pl, const float *sp, const float *ep,\n'
                   '                    const double *old_ep, double md, double *out) {\n'
                   '    plpslam_b200::trimming_camera c{cam4[0], cam4[1], cam4[2], cam4[3]};\n'
                   '    return plpslam_b200::endpoint_trimming(c, pose, pl, sp[0], sp[1], ep[0], ep[1], old_ep, md, out) ? 1 : 0;\n}\n')
    so = d / "libtrim.so"
    res = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                          f"-I{ROOT / 'structure-plp-slam_b200' / 'host'}", str(src), "-o", str(so)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return C.CDLL(str(so))


def _np_trim(cam4, T, L, sp, ep, old, md):
    fx, fy, cx, cy = cam4
    R, t = T[:3, :3], T[:3, 3]
    sk = lambda w: np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])
    Kl = np.array([[fy, 0, 0], [0, fx, 0], [-fy * cx, -fx * cy, fx * fy]])
    H = np.zeros((6, 6))
    H[:3, :3], H[3:, 3:], H[:3, 3:] = R, R, sk(t) @ R
    l1, l2, l3 = Kl @ (H @ L)[:3]
    P = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]]) @ T[:3]
    M = np.zeros((4, 4))
    M[:3, :3], M[:3, 3], M[3, :3] = sk(L[:3]), L[3:], -L[3:]
    out = []
    for (x, y) in (sp, ep):
        xc = -(y - (l2 / l1) * x + (l3 / l2)) * ((l1 * l2) / (l1 * l1 + l2 * l2))
        yc = -(l1 / l2) * xc - (l3 / l2)
        X = M @ (P.T @ np.cross([xc, yc, 1.0], [0.0, y - (l2 / l1) * x, 1.0]))
        out.append(X[:3] / X[3])
    out = np.concatenate(out)
    keep = not (np.linalg.norm(out[:3] - old[:3]) / md > 0.1 or np.linalg.norm(out[3:] - old[3:]) / md > 0.1)
    return keep, out


def _call(fn, cam4, T, L, sp, ep, old, md):
    out = np.zeros(6)
    d = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(_P)
    f = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(_P)
    keep = fn(d(cam4), d(T), d(L), f(sp), f(ep), d(old), C.c_double(md), out.ctypes.data_as(_P))
    return bool(keep), out


def test_endpoint_trimming_matches_oracle_and_numpy_and_recovers_the_segment(trim_lib, orc):
    rng = np.random.default_rng(0)
    cam4 = np.array([synth.FX, synth.FY, synth.CX, synth.CY])
    kept = erased = 0
    for k in range(200):
        T = np.eye(4)
        T[:3, :3] = synth.so3_exp(rng.normal(0, 0.2, 3))
        T[:3, 3] = rng.normal(0, 0.3, 3)
        Pc, Qc = rng.uniform([-2, -1.5, 3], [2, 1.5, 9]), rng.uniform([-2, -1.5, 3], [2, 1.5, 9])
        Pw, Qw = T[:3, :3].T @ (Pc - T[:3, 3]), T[:3, :3].T @ (Qc - T[:3, 3])
        L = synth.plucker_from_endpoints(Pw[None], Qw[None])[0]
        sp = np.array([synth.FX * Pc[0] / Pc[2] + synth.CX, synth.FY * Pc[1] / Pc[2] + synth.CY])
        ep = np.array([synth.FX * Qc[0] / Qc[2] + synth.CX, synth.FY * Qc[1] / Qc[2] + synth.CY])
        noise = rng.normal(0, 1.0 if k % 4 else 40.0, 4)
        sp_n, ep_n = (sp + noise[:2]).astype(np.float32), (ep + noise[2:]).astype(np.float32)
        old = np.concatenate([Pw, Qw]) + rng.normal(0, 0.02, 6)
        md = 6.0
        trim_lib.trim.restype = C.c_int
        orc.lib.orc_endpoint_trimming.restype = C.c_int
        g_keep, g_out = _call(trim_lib.trim, cam4, T, L, sp_n, ep_n, old, md)
        o_keep, o_out = _call(orc.lib.orc_endpoint_trimming, cam4, T, L, sp_n, ep_n, old, md)
        n_keep, n_out = _np_trim(cam4, T, L, sp_n.astype(np.float64), ep_n.astype(np.float64), old, md)
        assert g_keep == o_keep == n_keep
        assert np.allclose(g_out, o_out, rtol=1e-10, atol=1e-9) and np.allclose(g_out, n_out, rtol=1e-8, atol=1e-7)
        kept += g_keep
        erased += not g_keep
        if k % 4:  # exact geometry check with noise-free end points: the original 3-D segment comes back
            _, exact = _call(trim_lib.trim, cam4, T, L, sp.astype(np.float32), ep.astype(np.float32), old, md)
            assert np.allclose(exact, np.concatenate([Pw, Qw]), atol=2e-3)
    assert kept > 100 and erased > 10
