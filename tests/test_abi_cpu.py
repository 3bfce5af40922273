"""CPU: the C-ABI library loads and exports every symbol include/graspldm_b200.h declares; no compute calls."""
import ctypes
import os
import re

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "graspldm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gldm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from graspldm_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/graspldm_b200.h but not exported"


def test_python_binding_table_matches_header():
    from graspldm_b200 import _lib
    assert sorted(_lib.exported_symbols()) == header_symbols()
    assert _lib.lib().gldm_version() >= 100


def test_bad_arguments_fail_without_touching_the_gpu():
    from graspldm_b200 import _lib
    import pytest
    with pytest.raises(RuntimeError, match="null pointer"):
        _lib.call("gldm_furthest_point_sampling", None, 1, 8, 2, None, None)
    with pytest.raises(RuntimeError, match="bad sizes"):
        _lib.call("gldm_ball_query", 16, 16, 1, 0, 4, 0.1, 4, 16, None)


def test_no_cpu_fallback():
    import pytest
    import torch
    from graspldm_b200 import _pvcnn_backend, engine
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _pvcnn_backend.avg_voxelize_forward(torch.zeros(1, 3, 8), torch.zeros(1, 3, 8, dtype=torch.int32), 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.encoder_forward(None, torch.zeros(1, 8, 3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "graspldm_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{fn} imports the oracle"
