"""Drop-in for geotransformer/modules/sinkhorn/learnable_sinkhorn.py on the CUDA path (same class name, constructor,
forward signature and state_dict key `alpha`)."""
import ctypes

import torch
import torch.nn as nn

from .. import _lib


def log_optimal_transport(scores, alpha, num_iterations, row_masks=None, col_masks=None):
    """scores (B, M, N) fp32 on the GPU, alpha 0-d / 1-element fp32 tensor on the GPU, masks (B, M) / (B, N) bool or None
    -> (B, M + 1, N + 1) fp32 (se3et_log_optimal_transport)."""
    _lib.require_cuda(scores, alpha, row_masks, col_masks)
    scores = scores.float().contiguous()
    b, m, n = scores.shape
    rm = None if row_masks is None else row_masks.to(torch.uint8).contiguous()
    cm = None if col_masks is None else col_masks.to(torch.uint8).contiguous()
    a = alpha.detach().float().reshape(1).contiguous()
    out = torch.empty((b, m + 1, n + 1), dtype=torch.float32, device=scores.device)
    _lib.check(_lib.lib().se3et_log_optimal_transport(
        _lib.ptr(scores), _lib.ptr(rm), _lib.ptr(cm), _lib.ptr(a), _lib.i64(b), _lib.i64(m), _lib.i64(n),
        _lib.i64(num_iterations), _lib.ptr(out), _lib.stream_ptr()), "log_optimal_transport")
    return out


class LearnableLogOptimalTransport(nn.Module):
    """learnable_sinkhorn.py:5-70 (inference: no autograd through the kernel)."""

    def __init__(self, num_iterations, inf=1e12):
        super().__init__()
        if inf != 1e12:
            raise NotImplementedError("CUDA path: inf = 1e12 (the reference default)")
        self.num_iterations = num_iterations
        self.register_parameter('alpha', torch.nn.Parameter(torch.tensor(1.0)))
        self.inf = inf

    @torch.no_grad()
    def forward(self, scores, row_masks=None, col_masks=None):
        return log_optimal_transport(scores, self.alpha, self.num_iterations, row_masks, col_masks)

    def __repr__(self):
        return self.__class__.__name__ + '(num_iterations={})'.format(self.num_iterations)
