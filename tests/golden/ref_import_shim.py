"""Generation-time only (needs /root/reference): makes the UNMODIFIED reference python modules importable
on a CPU-only host so that golden vectors can be produced from them.

Nothing here is reference code: it installs stand-ins for third-party packages that are absent from this
image (trimesh, e3nn, easydict, open3d, IPython, ...), neutralises `.cuda()` and the mkdirs in config.py, and
provides `geotransformer.ext` from the reference build in oracle/_ref.  Import this module BEFORE anything
from `geotransformer`.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    return m


class _Trimesh:
    """Just enough of trimesh.base.Trimesh for utils_epn/rotation.py (octa/tetra/icosahedron constants)."""

    def __init__(self, vertices, faces):
        self.vertices = np.asarray(vertices, dtype=np.float64)
        self.faces = np.asarray(faces, dtype=np.int64).copy()

    def _normals(self):
        v = self.vertices
        f = self.faces
        n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
        return n / np.linalg.norm(n, axis=1, keepdims=True)

    def fix_normals(self):
        centroid = self.vertices.mean(0)
        n = self._normals()
        fc = self.vertices[self.faces].mean(1)
        flip = np.einsum("ij,ij->i", n, fc - centroid) < 0
        self.faces[flip] = self.faces[flip][:, ::-1]

    @property
    def face_normals(self):
        return self._normals()

    @property
    def edges(self):
        f = self.faces
        return np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)

    @property
    def edges_unique(self):
        e = np.sort(self.edges, axis=1)
        return np.unique(e, axis=0)

    @property
    def vertex_neighbors(self):
        nb = [[] for _ in range(len(self.vertices))]
        for a, b in self.edges_unique:
            nb[a].append(int(b))
            nb[b].append(int(a))
        return np.array([sorted(x) for x in nb])

    @property
    def face_adjacency(self):
        owner = {}
        pairs = []
        for fi, f in enumerate(self.faces):
            for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
                key = (min(a, b), max(a, b))
                if key in owner:
                    pairs.append((owner[key], fi))
                else:
                    owner[key] = fi
        return np.array(sorted(pairs))


def _icosahedron(_path=None):
    t = (1.0 + 5 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]])
    return _Trimesh(v, f)


class _EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


class _EmptyFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Empty modules for the compiled vgtk CUDA extensions (imported, never executed on the SE3ET path)."""
    NAMES = {"vgtk.cuda.zpconv", "vgtk.cuda.gathering", "vgtk.cuda.grouping"}

    def find_spec(self, fullname, path, target=None):
        if fullname in self.NAMES:
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        return types.ModuleType(spec.name)

    def exec_module(self, module):
        pass


def _ext_module():
    from oracle import points as op

    def radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius):
        nb = op.ref_radius_neighbors_raw(q_points.numpy(), s_points.numpy(), q_lengths.numpy(), s_lengths.numpy(),
                                         radius)
        return torch.from_numpy(nb)

    def grid_subsampling(points, lengths, normals, voxel_size):
        sp, sl, sn = op.ref_grid_subsampling_raw(points.numpy(), lengths.numpy(), normals.numpy(), voxel_size)
        return [torch.from_numpy(sp), torch.from_numpy(sl), torch.from_numpy(sn)]

    return _module("geotransformer.ext", radius_neighbors=radius_neighbors, grid_subsampling=grid_subsampling)


_installed = False


def install(variant="se3eti.3dmatch"):
    global _installed
    if _installed:
        return
    _installed = True
    for p in (REF, os.path.join(REF, "geotransformer/modules/e2pn/vgtk"), os.path.join(REF, "experiments", variant)):
        if p not in sys.path:
            sys.path.insert(0, p)
    tm = _module("trimesh", load=_icosahedron)
    tm.base = _module("trimesh.base", Trimesh=_Trimesh)
    _module("IPython", embed=lambda *a, **k: None)
    e3 = _module("e3nn")
    e3.o3 = _module("e3nn.o3")
    _module("easydict", EasyDict=_EasyDict)
    o3d = _module("open3d")
    for sub in ("geometry", "utility", "visualization", "pipelines", "io"):
        setattr(o3d, sub, _module("open3d." + sub))
    mpl = _module("matplotlib")
    mpl.pyplot = _module("matplotlib.pyplot")
    mpl.colors = _module("matplotlib.colors")
    mpl.cm = _module("matplotlib.cm")
    mpl.use = lambda *a, **k: None
    _module("mpl_toolkits")
    _module("mpl_toolkits.mplot3d", Axes3D=object)
    _module("turtle", forward=lambda *a, **k: None)
    _module("plyfile", PlyData=object, PlyElement=object)
    _module("colour", Color=object)
    _module("parse", parse=lambda *a, **k: None)
    _module("ipdb", set_trace=lambda *a, **k: None)
    _module("coloredlogs", install=lambda *a, **k: None, ColoredFormatter=object)
    _module("tensorboardX", SummaryWriter=object)
    sys.meta_path.insert(0, _EmptyFinder())
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.empty_cache = lambda: None
    import geotransformer  # noqa: F401  (the package itself is plain python)
    _ext_module()
    import geotransformer.utils.common as common
    common.ensure_dir = lambda p: None
    # import order matters (circular import in the reference)
    import geotransformer.modules.geotransformer  # noqa: F401
    import geotransformer.utils.data as data
    data.estimate_normals = lambda p: np.zeros((len(p), 3))


def make_cfg(variant="se3eti.3dmatch"):
    install(variant)
    import importlib
    cfg_mod = importlib.import_module("config")
    return cfg_mod.make_cfg()
