"""The reference-side adapter header (structure-plp-slam_b200/host/plpslam_b200_adapter.hpp) is compiled INSIDE the
reference tree, whose dependencies (OpenCV, Eigen, DBoW2) are not installed here.  This test compiles it with
`g++ -fsyntax-only` against tests/adapter_mock/ -- mock declarations carrying the reference's member names and types --
so that typos and drift between the adapters and include/plpslam_b200.h are caught on the CPU."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_adapter_header_compiles_against_mocks(tmp_path):
    tu = tmp_path / "adapter_tu.cc"
    tu.write_text('#include "plpslam_b200_adapter.hpp"\nint main() { return 0; }\n')
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-DPLPSLAM_B200_WITH_REFERENCE_TYPES", "-DUSE_DBOW2",
           f"-I{ROOT / 'tests' / 'adapter_mock'}", f"-I{ROOT / 'include'}",
           f"-I{ROOT / 'structure-plp-slam_b200' / 'host'}", str(tu)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[:4000]
