"""The plane RANSAC DEVICE code (structure-plp-slam_b200/csrc/plane_kernels.cuh) executed on the CPU: tests/cta_emu compiles
the same kernel text for the host (one host thread per CUDA thread, a pthread barrier for __syncthreads, blocks one
after the other) and this test compares it with the oracle bit for bit.  It checks the kernels' logic -- indexing, phase
structure, replay bookkeeping -- in a container without a GPU; the GPU parity run is tests/test_zz_plane_gpu.py."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

import plane_data

ROOT = Path(__file__).resolve().parent.parent
_P = C.c_void_p


class Cfg(C.Structure):
    _fields_ = [("mode", C.c_int32), ("points_per_ransac", C.c_int32), ("planar_distance_thresh", C.c_double),
                ("final_error_thresh", C.c_double), ("inliers_ratio_thr", C.c_double), ("initial_best_error", C.c_double)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("emu") / "libplane_emu.so"
    cmd = ["g++", "-O1", "-std=c++17", "-pthread", "-shared", "-fPIC", "-ffp-contract=off",
           f"-I{ROOT / 'structure-plp-slam_b200' / 'csrc'}", f"-I{ROOT / 'tests' / 'cta_emu'}",
           str(ROOT / "tests" / "cta_emu" / "plane_emu.cc"), "-o", str(so)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[:3000]
    return C.CDLL(str(so))


def _run(emu, pts, valid, smp, cfg, eq0=(0, 0, 0, 0), err0=0.0):
    P = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
    sm = np.ascontiguousarray(smp, np.int32)
    v = None if valid is None else np.ascontiguousarray(valid, np.uint8)
    c = Cfg(cfg["mode"], cfg["points_per_ransac"], cfg["planar_distance_thresh"], cfg["final_error_thresh"],
            cfg["inliers_ratio_thr"], cfg.get("initial_best_error", 0.0))
    eq = np.array(eq0, np.float64)
    err = C.c_double(err0)
    inl = np.zeros(max(len(P), 1), np.uint8)
    st = emu.emu_plane_ransac(P.ctypes.data_as(_P), None if v is None else v.ctypes.data_as(_P), C.c_int(len(P)),
                              sm.ctypes.data_as(_P), C.c_int(sm.shape[0]), C.c_int(sm.shape[1]), C.byref(c),
                              eq.ctypes.data_as(_P), C.byref(err), inl.ctypes.data_as(_P))
    return int(st), eq, float(err.value), inl[:len(P)].copy()


@pytest.mark.parametrize("seed", range(4))
def test_plane_kernels_on_cpu_match_oracle(emu, orc, seed):
    n = [300, 120, 60, 18][seed]
    pts, valid, _, _ = plane_data.make_plane_cloud(seed + 10, n=n)
    smp = plane_data.draw_plane_samples(seed, valid, 12, 18)
    want = orc.plane_ransac(pts, valid, smp, plane_data.CFG_ESTIMATE)
    got = _run(emu, pts, valid, smp, plane_data.CFG_ESTIMATE)
    assert got[0] == want[0] and np.array_equal(got[1], want[1]) and got[2] == want[2] and np.array_equal(got[3], want[3])
    # update mode from a stored plane
    eq0, err0 = orc.plane_fit(pts, np.nonzero(valid)[0][:18].astype(np.int32))
    smp_u = plane_data.draw_plane_samples(seed + 1, valid, 6, int(np.ceil(0.8 * n)))
    want = orc.plane_ransac(pts, valid, smp_u, plane_data.CFG_UPDATE, eq0, err0)
    got = _run(emu, pts, valid, smp_u, plane_data.CFG_UPDATE, eq0, err0)
    assert got[0] == want[0] and np.array_equal(got[1], want[1]) and got[2] == want[2] and np.array_equal(got[3], want[3])


def test_plane_kernels_on_cpu_failure_paths(emu, orc):
    cloud = np.random.default_rng(1).uniform(-1, 1, (80, 3))
    smp = plane_data.draw_plane_samples(2, np.ones(80), 8, 18)
    want = orc.plane_ransac(cloud, None, smp, plane_data.CFG_ESTIMATE)
    got = _run(emu, cloud, None, smp, plane_data.CFG_ESTIMATE)
    assert got[0] == want[0] == 0 and np.array_equal(got[1], want[1]) and got[2] == want[2] and got[3].sum() == 0


# ---------------------------------------------------------------------------------------------------------------------
# the essential-matrix RANSAC kernels are GPU-verified (tests/test_essential_gpu.py): running them through the emulator as
# well validates the emulator itself
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ess_emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("emu") / "libess_emu.so"
    cmd = ["g++", "-O1", "-std=c++17", "-pthread", "-shared", "-fPIC", "-ffp-contract=off",
           f"-I{ROOT / 'structure-plp-slam_b200' / 'csrc'}", f"-I{ROOT / 'tests' / 'cta_emu'}",
           str(ROOT / "tests" / "cta_emu" / "essential_emu.cc"), "-o", str(so)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[:3000]
    return C.CDLL(str(so))


@pytest.mark.parametrize("seed", range(3))
def test_essential_kernels_on_cpu_match_oracle(ess_emu, orc, seed):
    import ess_data
    n = [200, 60, 33][seed]
    b1, b2, matches, _ = ess_data.make_two_view(seed + 10, n=n)
    smp = ess_data.draw_samples(seed, len(matches), 12)
    for recompute in (False, True):
        want = orc.essential_ransac(b1, b2, matches, smp, recompute)
        B1, B2 = np.ascontiguousarray(b1, np.float64), np.ascontiguousarray(b2, np.float64)
        m, sm = np.ascontiguousarray(matches, np.int32), np.ascontiguousarray(smp, np.int32)
        inl, E, score = np.zeros(len(m), np.uint8), np.zeros(9), C.c_double(0)
        valid = ess_emu.emu_essential_ransac(B1.ctypes.data_as(_P), B2.ctypes.data_as(_P), m.ctypes.data_as(_P),
                                             C.c_int(len(m)), sm.ctypes.data_as(_P), C.c_int(len(sm)),
                                             C.c_int(1 if recompute else 0), inl.ctypes.data_as(_P), E.ctypes.data_as(_P),
                                             C.byref(score))
        assert valid == want[0] and np.array_equal(inl, want[1]) and np.array_equal(E.reshape(3, 3), want[2])
        assert score.value == want[3]
