"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/b2g.h declares
(no compute calls here -- there is no GPU in the build container)."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "multi-modal-gnn_b200"


def _declared():
    text = open(os.path.join(ROOT, "include", "b2g.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2g_\w+)\s*\(", text)))


def test_library_builds_and_exports_all_declared_symbols():
    L = importlib.import_module(PKG + "._lib")
    path = L.build()
    assert os.path.isfile(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b2g.h but not exported by libb2g.so"
    bound = set(L.exported_symbols())
    assert set(names) == bound, (sorted(set(names) - bound), sorted(bound - set(names)))
    assert L.load().b2g_version() >= 100
    assert L.load().b2g_launch_count() == 0


def test_library_is_sm100a_only():
    L = importlib.import_module(PKG + "._lib")
    assert any("compute_100a" in f for f in L.NVCC_FLAGS) and "-lineinfo" in L.NVCC_FLAGS


def test_cpu_tensors_are_rejected_loudly():
    import torch
    ops = importlib.import_module(PKG + ".ops")
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.linear(torch.randn(4, 8), torch.randn(3, 8), None)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, PKG)):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), f
