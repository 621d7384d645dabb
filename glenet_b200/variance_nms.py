"""GLENet's variance-voting NMS (``new_nms_gpu`` / ``nms_func``) and soft-NMS (``softnms_gpu`` / ``softnms``) on the device.

The reference implements these as Python loops in ``pcdet/ops/iou3d_nms/iou3d_nms_utils.py`` (:200-273, :292-356): an
N x N ``boxes_bev_iou_cpu`` on one CPU core (3.4 s for N = 4096), then one numpy / torch iteration per retired box.  Here
the matrix comes from this library's IoU kernel and STAYS on the GPU, and the whole loop is one kernel
(``csrc/vnms.cu``, ``glenet_variance_nms_gpu``): per iteration an arg max, one column of the matrix, the voters compacted
in index order and the reference's float32 sums run sequentially in that order (numpy's reduction order), the score
update.  What crosses PCIe is the boxes and scores (28 N + 4 N bytes each way), not the matrix (67 MB at N = 4096).

Same names, signatures and return conventions as the reference: ``new_nms_gpu`` returns ``(keep [numpy], None, boxes
[numpy])``, ``softnms_gpu`` returns tensors.  ``new_nms_gpu`` uses the CPU dialect of the IoU (the reference calls
``boxes_bev_iou_cpu``): host-libm trigonometry per box, no FMA contraction.

One deliberate shortcut in ``nms_func``, as in round 1: the reference keeps iterating over boxes whose score has already
been multiplied to 0 (``score_threshold`` defaults to 0 and ``0 < 0`` is false), voting new coordinates for boxes that can
never be kept.  Those iterations change neither ``keep`` nor the rows ``new_boxes[keep]`` the caller reads
(``model_nms_utils.py:44-45``), so the device loop stops at the last positive score; rows of suppressed boxes keep
their input values.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib

_MODE = {"hard": 0, "gaussian": 1, "linear": 2}


def _limit_period(val: torch.Tensor, offset: float = 0.5, period: float = math.pi) -> torch.Tensor:
    """pcdet/utils/common_utils.py:21-24 on a float32 CPU tensor (the reference routes numpy input through torch as well)."""
    return val - torch.floor(val / period + offset) * period


def _device_loop(boxes_d, scores_d, variance_d, iou_d, iou_threshold, score_threshold, mode, soft_sigma=0.3):
    """Run csrc/vnms.cu in place on (n, 7) boxes / (n,) scores with the resident (n, n) IoU matrix."""
    n = boxes_d.shape[0]
    if n == 0:
        return
    lib = _lib.load()
    dev = boxes_d.device
    assert boxes_d.is_contiguous() and scores_d.is_contiguous() and iou_d.is_contiguous() and iou_d.shape == (n, n)
    var_ptr, var_cols = None, 0
    if variance_d is not None:
        variance_d = variance_d.contiguous()
        var_ptr, var_cols = variance_d.data_ptr(), variance_d.shape[1]
    with torch.cuda.device(dev):
        rc = lib.glenet_variance_nms_gpu(boxes_d.data_ptr(), scores_d.data_ptr(), var_ptr, var_cols, iou_d.data_ptr(), 1, n,
                                         float(iou_threshold), float(score_threshold), _MODE[mode], float(soft_sigma),
                                         torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "glenet_variance_nms_gpu")


def nms_func(boxes, scores, iou_threshold, score_threshold=0, variance=None):
    """iou3d_nms_utils.py:227-273.  ``boxes`` (N, 7) and ``scores`` (N,) are float32 numpy arrays that are updated in
    place, as in the reference; ``variance`` (N, >= 7) or None.  Returns ``(scores, boxes)``."""
    from .iou3d_nms_utils import _bev_iou_cpu_dialect_on_device
    b_h = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float32))
    if b_h.shape[0] == 0:
        return scores, boxes
    iou_d = _bev_iou_cpu_dialect_on_device(b_h, b_h)                     # (N, N), stays on the GPU
    dev = iou_d.device
    b_d = b_h.to(dev)
    s_d = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32)).to(dev)
    v_d = None if variance is None else torch.from_numpy(np.ascontiguousarray(variance, dtype=np.float32)).to(dev)
    _device_loop(b_d, s_d, v_d, iou_d, iou_threshold, score_threshold, "hard")
    boxes[...] = b_d.cpu().numpy()
    scores[...] = s_d.cpu().numpy()
    return scores, boxes


def new_nms_gpu(boxes, scores, iou_threshold, pre_maxsize=None, score_threshold=0, variance=None, **kwargs):
    """
    :param boxes: (N, 7) [x, y, z, dx, dy, dz, heading]
    :param scores: (N)
    :param thresh:
    :return: (keep indices sorted by descending new score [numpy], None, voted boxes [numpy])

    iou3d_nms_utils.py:200-224.  ``**kwargs`` swallows the NMS_CONFIG dict (model_nms_utils.py:40-43); ``pre_maxsize`` is
    accepted and ignored, as there.
    """
    boxes_h = boxes.detach().cpu().float().numpy().copy()
    scores_h = scores.detach().cpu().float().numpy().copy()
    variance_h = variance.detach().cpu().float().numpy() if variance is not None else None
    boxes_h[:, 6] = _limit_period(torch.from_numpy(boxes_h[:, 6].copy()), offset=0.5, period=math.pi * 2).numpy()
    new_scores, new_boxes = nms_func(boxes_h, scores_h, iou_threshold, score_threshold, variance=variance_h)
    keep = np.flatnonzero(new_scores > 0)
    keep = keep[np.argsort(new_scores[keep])[::-1]]
    return keep, None, new_boxes


def scale_by_iou(ious, soft_sigma, soft_mode="gaussian"):
    """Score decay of soft-NMS (iou3d_nms_utils.py:303-310): ``1 - iou`` where ``iou >= soft_sigma`` (linear), or
    ``exp(-iou^2 / soft_sigma)`` (gaussian)."""
    if soft_mode == "linear":
        return torch.where(ious >= soft_sigma, 1 - ious, torch.ones_like(ious))
    return torch.exp(-ious ** 2 / soft_sigma)


def softnms(boxes, scores, iou_threshold, soft_sigma, score_threshold, soft_mode="gaussian", variance=None):
    """iou3d_nms_utils.py:312-356 on CUDA tensors, in place.  The reference launches ``boxes_iou_bev(remaining, top)`` in
    every iteration; all of those IoUs are between original boxes, so one N x N launch and the device loop replace them."""
    from .iou3d_nms_utils import boxes_iou_bev
    assert soft_mode in ["linear", "gaussian"]
    if not (boxes.is_cuda and scores.is_cuda):
        raise RuntimeError("softnms expects CUDA tensors")
    b = boxes[:, :7].contiguous().float()
    iou_d = boxes_iou_bev(b, b)
    s = scores.contiguous().float()
    _device_loop(b, s, None if variance is None else variance.float(), iou_d, iou_threshold, score_threshold, soft_mode, soft_sigma)
    boxes[:, :7] = b
    scores.copy_(s)
    return scores, boxes


def softnms_gpu(boxes, scores, iou_threshold, score_threshold=0.1, soft_mode='gaussian', variance=None, soft_sigma=0.3, **kwargs):
    """iou3d_nms_utils.py:292-301: ``(keep sorted by descending decayed score, None, voted boxes)``, all CUDA tensors."""
    assert soft_mode in ["linear", "gaussian"]
    assert boxes.shape[-1] == 7
    new_scores, new_boxes = softnms(boxes, scores, iou_threshold, soft_sigma, score_threshold, soft_mode, variance=variance)
    keep = torch.nonzero(new_scores > score_threshold, as_tuple=False).view(-1)
    keep = keep[torch.argsort(new_scores[keep], descending=True)]
    return keep, None, new_boxes
