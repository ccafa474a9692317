"""Builds se3et_b200/csrc/libse3et_b200.so (the C-ABI library) with nvcc for sm_100a, in-tree.

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import glob
import hashlib
import os
import subprocess

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(_CSRC, "libse3et_b200.so")
_STAMP = os.path.join(_CSRC, ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _sources():
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    for p in _sources() + sorted(glob.glob(os.path.join(_CSRC, "*.cuh"))) + sorted(
            glob.glob(os.path.join(_CSRC, "..", "..", "include", "*.h"))):
        h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library. Rebuilds only when sources changed."""
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(_STAMP) and open(_STAMP).read() == digest:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in _sources():
        obj = src[:-3] + ".o"
        objs.append(obj)
        procs.append((src, subprocess.Popen([nvcc, *NVCC_FLAGS, "-c", src, "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out))
    subprocess.check_call([nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB_PATH, *objs, "-lcudart"])
    with open(os.path.join(_CSRC, "build.log"), "w") as f:
        f.write("\n".join(log))
    with open(_STAMP, "w") as f:
        f.write(digest)
    if verbose:
        print("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
