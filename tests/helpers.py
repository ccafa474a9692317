"""Shared test helpers: deterministic parameters (so fixtures need not store weights) and small model configs."""
import zlib

import numpy as np
import torch

CONST_KEYS = ("kernel_points", "anchors", "quotient_anchors", "kidx_rot", "ridx_rot", "div_term", "trace_idx")

SMALL_CFG = dict(
    init_dim=16, group_norm=4, output_dim=32, input_dim=1, init_voxel=0.025, init_radius=0.0625, init_sigma=0.05,
    hidden_dim=64, num_heads=4, tr_output_dim=64, sigma_d=0.2, sigma_a=15.0, angle_k=3,
    blocks=['self_eq', 'cross', 'self_eq', 'cross', 'self_eq', 'cross'],
)


def is_constant(name):
    return any(name.endswith(k) or (k in name.split(".")[-1]) for k in CONST_KEYS)


def seeded_tensor(name, shape):
    """Deterministic fp32 values for parameter `name` (CPU generator seeded by crc32(name))."""
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    shape = tuple(shape)
    leaf = name.split(".")[-1]
    if leaf == "weights":  # KPConvInterSO3 (K_real, A, Cin, Cout)
        fan_in = 15 * shape[1] * shape[2] / 2.5
        return torch.randn(shape, generator=g) * (1.5 / fan_in ** 0.5)
    if leaf == "weight" and len(shape) == 2:
        return torch.randn(shape, generator=g) * (1.0 / shape[1] ** 0.5)
    if leaf == "weight":  # norm scale
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf == "bias":
        return 0.1 * torch.randn(shape, generator=g)
    return torch.randn(shape, generator=g)


def seeded_state_dict(template):
    """template: mapping name -> tensor (only shapes/dtypes used). Constants are passed through unchanged."""
    out = {}
    for name, t in template.items():
        if is_constant(name) or not torch.is_floating_point(t):
            out[name] = t.clone()
        else:
            out[name] = seeded_tensor(name, t.shape)
    return out


def small_pair(index=13, crop=0.9):
    """A dense ~0.9 m crop of a synthetic 3DMatch-shaped pair: ~2.6k points, 58 superpoints."""
    from se3et_b200 import synthetic
    p = synthetic.make_3dmatch_pair(index, crop=crop)
    pts = np.concatenate([p["ref_points"], p["src_points"]])
    lens = np.array([len(p["ref_points"]), len(p["src_points"])])
    return pts, lens
