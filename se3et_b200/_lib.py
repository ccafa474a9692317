"""ctypes binding of the C-ABI library (include/se3et_b200.h).

There is no CPU fallback: importing works anywhere (so the CPU test suite can check the
exported symbols), but every compute call requires CUDA tensors and raises if the library is
missing or a call fails.
"""
import ctypes
import os
import re
import threading

import torch

from . import build as _build

_HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "se3et_b200.h")

SE3ET_STATUS_WORDS = 8
STATUS_ERROR, STATUS_M_TOTAL, STATUS_MAX_COUNT, STATUS_REQ_KCELLS, STATUS_MAX_LENGTH = 0, 1, 2, 3, 4
DEV_GRID_TOO_LARGE, DEV_INDEX_RANGE = 1, 2

_ERRORS = {-1: "CUDA error", -2: "invalid argument", -3: "workspace too small", -4: "unsupported configuration"}

_lib = None

# kernels launched by one call of each entry point (cudaMemsetAsync nodes are not counted)
KERNELS_PER_CALL = {
    "se3et_grid_subsample": 16, "se3et_radius_neighbors": 9, "se3et_gemm_bf16": 1, "se3et_gemm_bf16_gnstats": 1, "se3et_gemm_bf16_gnapply": 2, "se3et_gemm_bf16_gnapply_dual": 2, "se3et_linear_gnstats_gram": 2, "se3et_linear_gnstats_gram2": 3, "se3et_linear_gnstats_stream": 1, "se3et_gemm_grouped_bf16": 1,
    "se3et_kpconv_gather": 1, "se3et_kpconv_fused": 1, "se3et_kpconv_rows": 1, "se3et_kpconv_cin1": 1, "se3et_kpconv_lift": 2, "se3et_groupnorm_stats": 1, "se3et_groupnorm_apply": 1, "se3et_groupnorm_double": 1, "se3et_maxpool_nbr": 1,
    "se3et_anchor_max": 1, "se3et_upsample_concat": 1, "se3et_geo_embed_indices": 1, "se3et_geo_embed_project": 1, "se3et_geo_embed_lookup": 1,
    "se3et_flash_attention": 1, "se3et_add_layernorm": 1, "se3et_linear_add_layernorm": 1, "se3et_l2_normalize_rows": 1,
    "se3et_superpoint_matching": 4, "se3et_anchor_pair_stats": 1, "se3et_anchor_mix_weights": 1,
    "se3et_anchor_mix": 1, "se3et_sh_bias_add": 1, "se3et_point_to_node_partition": 5, "se3et_log_optimal_transport": 1, "se3et_lgr_correspondences": 1, "se3et_lgr_register": 2, "se3et_weighted_procrustes": 1,
}


class _Instrumented:
    """Thin proxy over the ctypes library: counts calls per entry point and, for the names in `timed`, brackets the
    call with CUDA events on the current stream (bench.py reads both; nothing is recorded unless enabled)."""

    def __init__(self, cdll):
        self._cdll = cdll
        self.counts = {}
        self.timed = set()
        self.events = {}
        self.enabled = False
        self._lock = threading.Lock()  # launch sequences may run on several host threads

    def __getattr__(self, name):
        # called once per entry point: the bound callable is stored on the instance, later lookups never get here
        fn = getattr(self._cdll, name)
        if not name.startswith("se3et_"):
            setattr(self, name, fn)
            return fn

        def call(*args):
            if not self.enabled:
                return fn(*args)
            with self._lock:
                self.counts[name] = self.counts.get(name, 0) + 1
            if name in self.timed:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                rc = fn(*args)
                b.record()
                with self._lock:
                    self.events.setdefault(name, []).append((a, b))
                return rc
            return fn(*args)

        call.restype = getattr(fn, "restype", None)
        setattr(self, name, call)
        return call

    def reset(self, timed=()):
        self.counts, self.events, self.timed = {}, {}, set(timed)

    def launches(self):
        return sum(KERNELS_PER_CALL.get(k, 1) * v for k, v in self.counts.items())

    def timed_ms(self, name):
        ev = self.events.get(name, [])
        return sum(a.elapsed_time(b) for a, b in ev), len(ev)


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
        # SE3ET_NO_REBUILD=1: load the library as it is (A/B measurements of prebuilt variants).  Otherwise a stale
        # binary is never loaded after a failed build: the C ABI may have changed under it.
        if os.path.exists("/usr/local/cuda/bin/nvcc") and os.environ.get("SE3ET_NO_REBUILD") != "1":
            path = _build.build_library()
        if not os.path.exists(path):
            raise RuntimeError(
                "se3et_b200: %s is missing and could not be built; the hot path has no CPU fallback" % path)
        cdll = ctypes.CDLL(path)
        cdll.se3et_last_error.restype = ctypes.c_char_p
        _lib = _Instrumented(cdll)
        _lib._name = cdll._name
    return _lib


def check(rc, what):
    if rc != 0:
        detail = lib().se3et_last_error().decode() if rc == -1 else ""
        raise RuntimeError("se3et_b200.%s failed: %s %s" % (what, _ERRORS.get(rc, rc), detail))


def _raw_stream(device_index=None):
    """Current CUDA stream handle of a device as an int.  torch.cuda.current_stream() builds a Stream object through
    several Python layers: at ~375 C-ABI calls per launch sequence it was a quarter of the host time."""
    if device_index is None:
        device_index = torch._C._cuda_getDevice()
    return torch._C._cuda_getCurrentRawStream(device_index)


def stream_ptr():
    return ctypes.c_void_p(_raw_stream())


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
    """Grow-only scratch buffer handed to the C ABI, one per (device, CUDA stream): launch sequences running
    concurrently on different streams must not share scratch memory."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes, device):
        key = (device.type, device.index, _raw_stream(device.index))
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
            self._buf[key] = buf
        return buf


workspace = Workspace()
