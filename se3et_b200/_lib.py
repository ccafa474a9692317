"""ctypes binding of the C-ABI library (include/se3et_b200.h).

There is no CPU fallback: importing works anywhere (so the CPU test suite can check the
exported symbols), but every compute call requires CUDA tensors and raises if the library is
missing or a call fails.
"""
import ctypes
import os
import re

import torch

from . import build as _build

_HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "se3et_b200.h")

SE3ET_STATUS_WORDS = 8
STATUS_ERROR, STATUS_M_TOTAL, STATUS_MAX_COUNT, STATUS_REQ_KCELLS = 0, 1, 2, 3
DEV_GRID_TOO_LARGE, DEV_INDEX_RANGE = 1, 2

_ERRORS = {-1: "CUDA error", -2: "invalid argument", -3: "workspace too small", -4: "unsupported configuration"}

_lib = None


def declared_symbols():
    """Every function name declared in include/se3et_b200.h."""
    src = open(_HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(se3et_[a-z0-9_]+)\s*\(", src)))


def lib():
    """Loads (building if the sources changed and nvcc is present) the shared library."""
    global _lib
    if _lib is None:
        path = _build.LIB_PATH
        if os.path.exists("/usr/local/cuda/bin/nvcc"):
            try:
                path = _build.build_library()
            except Exception:
                if not os.path.exists(path):
                    raise
        if not os.path.exists(path):
            raise RuntimeError(
                "se3et_b200: %s is missing and could not be built; the hot path has no CPU fallback" % path)
        _lib = ctypes.CDLL(path)
        _lib.se3et_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc, what):
    if rc != 0:
        detail = lib().se3et_last_error().decode() if rc == -1 else ""
        raise RuntimeError("se3et_b200.%s failed: %s %s" % (what, _ERRORS.get(rc, rc), detail))


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def i64(v):
    return ctypes.c_int64(int(v))


def f32(v):
    return ctypes.c_float(float(v))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("se3et_b200: expected CUDA tensors (the hot path has no CPU fallback)")


class Workspace:
    """Grow-only per-device scratch buffer handed to the C ABI."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes, device):
        key = (device.type, device.index)
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
            self._buf[key] = buf
        return buf


workspace = Workspace()
