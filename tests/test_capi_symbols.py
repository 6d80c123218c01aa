"""CPU checks of the C-ABI boundary: the library loads, exports every declared symbol, and refuses
to compute without a GPU (no silent CPU fallback)."""
import ctypes as C

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(plp):
    lib = plp.lib()
    names = plp.declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/plpslam_b200.h but not exported: {missing}"


def test_no_cpu_fallback_without_device(plp):
    lib = plp.lib()
    if lib.plp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    st = lib.plp_ctx_create(C.c_int(0), C.byref(h))
    assert st == 2  # PLP_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.plp_last_error()
    with pytest.raises(plp.PlpError):
        plp.Context(0)


def test_product_never_imports_oracle():
    """The product tree must not reference oracle/ (it is test infrastructure)."""
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent / "structure-plp-slam_b200"
    bad = []
    for p in list(root.rglob("*.py")) + list(root.rglob("*.cu")) + list(root.rglob("*.cuh")) + list(root.rglob("*.hpp")):
        txt = p.read_text(errors="ignore")
        if "liboracle" in txt or "oracle/" in txt or "oracle_api" in txt or "import oracle" in txt:
            bad.append(str(p))
    assert not bad, bad
