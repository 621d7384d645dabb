"""Drop-in for ``iou3d`` of ``cvae_uncertainty/eval_utils/eval_utils.py:14-65`` -- the recall IoU of the CVAE evaluation.

The reference computes it per (ground truth, prediction) pair with Python loops over numpy float32 scalars on the host
(``pcdet/utils/loss_utils.py:276-411,551-635``) after ``.cpu().numpy()`` round trips; here it is one kernel launch
(``glenet_cvae_iou3d_gpu``, ``csrc/rotate_iou.cu``) in the same float32 dialect.  No CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["iou3d"]


def iou3d(gboxes, qboxes):
    '''
        gboxes / qboxes: [N, 7], [x, y, z, w, l, h, ry] in velo coord.
        Notice: (x, y, z) is the real center of bbox.

    Returns the (N,) float32 IoU of pair i (eval_utils.py:64-65); (0, 1) zeros for N == 0 (:26-27).
    '''
    assert gboxes.shape[0] == qboxes.shape[0]
    n = gboxes.shape[0]
    if n == 0:
        return torch.zeros((0, 1), device=gboxes.device, dtype=torch.float32)
    if not (gboxes.is_cuda and qboxes.is_cuda):
        raise RuntimeError("glenet_b200.cvae_eval_utils.iou3d expects CUDA tensors (eval_utils.py:217-219 moves them there); there is no CPU fallback")
    if gboxes.dtype != torch.float32 or qboxes.dtype != torch.float32:
        raise RuntimeError("iou3d expects float32 boxes")
    assert gboxes.dim() == 2 and gboxes.shape[1] == 7 and qboxes.shape == gboxes.shape
    dev = gboxes.device
    g, q = gboxes.contiguous(), qboxes.to(dev).contiguous()
    out = torch.empty((n,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().glenet_cvae_iou3d_gpu(g.data_ptr(), q.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "glenet_cvae_iou3d_gpu")
    return out
