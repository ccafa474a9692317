"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/se3et_b200.h declares.
No compute call is made here (there is no GPU in the build container)."""
import ctypes

import pytest
import torch

from se3et_b200 import _lib, ext


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert "se3et_grid_subsample" in names and "se3et_radius_neighbors" in names
    for n in names:
        assert hasattr(L, n), n
    assert L.se3et_version() >= 100


def test_argument_validation_without_gpu():
    L = _lib.lib()
    nbytes = ctypes.c_size_t(0)
    assert L.se3et_grid_subsample_workspace_bytes(_lib.i64(1000), _lib.i64(2), _lib.i64(1 << 20), ctypes.byref(nbytes)) == 0
    assert nbytes.value > 0
    assert L.se3et_grid_subsample_workspace_bytes(_lib.i64(-1), _lib.i64(2), _lib.i64(1 << 20), ctypes.byref(nbytes)) == -2
    assert L.se3et_radius_neighbors_workspace_bytes(_lib.i64(10), _lib.i64(10), _lib.i64(0), ctypes.byref(nbytes)) == -2


def test_no_cpu_fallback():
    pts = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ext.grid_subsampling(pts, torch.tensor([4]), pts, 0.1)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ext.radius_neighbors(pts, pts, torch.tensor([4]), torch.tensor([4]), 0.1)


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under se3et_b200/ may import or execute it (no CPU fallback)."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "se3et_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|importlib\.import_module\(['\"]oracle", re.M)
    bad = []
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                if pat.search(src):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
