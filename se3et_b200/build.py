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


def _local_includes(path, seen):
    """Transitive closure of the quoted #includes of `path` that resolve under csrc/ or include/."""
    import re
    for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', open(path).read(), flags=re.M):
        for base in (os.path.dirname(path), _CSRC, os.path.join(_CSRC, "..", "..", "include")):
            cand = os.path.normpath(os.path.join(base, inc))
            if os.path.exists(cand):
                if cand not in seen:
                    seen.add(cand)
                    _local_includes(cand, seen)
                break
    return seen


def _object_digest(src):
    h = hashlib.sha256()
    h.update(open(src, "rb").read())
    for dep in sorted(_local_includes(src, set())):
        h.update(open(dep, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Rebuilds only the objects whose source (or a header it
    includes) changed, then relinks."""
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(_STAMP) and open(_STAMP).read() == digest:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in _sources():
        obj = src[:-3] + ".o"
        objs.append(obj)
        od, ostamp = _object_digest(src), src[:-3] + ".o.stamp"
        if not force and os.path.exists(obj) and os.path.exists(ostamp) and open(ostamp).read() == od:
            continue
        procs.append((src, ostamp, od, subprocess.Popen([nvcc, *NVCC_FLAGS, "-c", src, "-o", obj],
                                                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = None
    for src, ostamp, od, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            if os.path.exists(ostamp):
                os.remove(ostamp)
            failed = failed or (src, out)
        else:
            with open(ostamp, "w") as f:
                f.write(od)
    if failed:
        raise RuntimeError("nvcc failed for %s:\n%s" % failed)
    subprocess.check_call([nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB_PATH, *objs, "-lcudart"])
    with open(os.path.join(_CSRC, "build.log"), "a" if len(procs) < len(objs) else "w") as f:
        f.write("\n".join(log))
    with open(_STAMP, "w") as f:
        f.write(digest)
    if verbose:
        print("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
