"""GLENet's variance-voting NMS and soft-NMS on top of the B200 IoU kernels.

The reference implements these in Python inside ``pcdet/ops/iou3d_nms/iou3d_nms_utils.py``:
``new_nms_gpu`` (:200-224) / ``nms_func`` (:227-273) -- the ``NMS_TYPE`` of every shipped GLENet
config (``tools/cfgs/kitti_models/GLENet_VR.yaml:178``) -- and ``softnms_gpu`` / ``scale_by_iou`` /
``softnms`` (:292-356).  They are host-side control flow around ``boxes_bev_iou_cpu`` (one N x N
matrix, 3.4 s on a CPU core for N = 4096) and ``boxes_iou_bev`` (one N x 1 launch per iteration).
Here the same control flow runs on top of the drop-in IoU functions, i.e. the N x N matrix is one
GPU launch.  Same signatures, same return conventions (``new_nms_gpu`` returns numpy, ``softnms_gpu``
returns tensors, both a 3-tuple ``(keep, None, new_boxes)``).

One deliberate shortcut in ``nms_func``: the reference keeps iterating over boxes whose score has
already been multiplied to 0 (``score_threshold`` defaults to 0 and ``0 < 0`` is false), voting new
coordinates for boxes that can never be kept.  Those iterations cannot change ``keep`` nor the rows
``new_boxes[keep]`` the caller reads (``model_nms_utils.py:44-45``), so the loop stops once every
remaining score is 0; rows of ``new_boxes`` for suppressed boxes therefore keep their input values.
"""
from __future__ import annotations

import numpy as np
import torch

_STD_IOU_SIGMA = 0.05   # iou3d_nms_utils.py:257,339


def _limit_period(val, offset=0.5, period=np.pi):
    """pcdet/utils/common_utils.py:21-24 for numpy input (computed through torch float32 like the reference)."""
    t = torch.from_numpy(val).float()
    return (t - torch.floor(t / period + offset) * period).numpy()


def nms_func(boxes, scores, iou_threshold, score_threshold=0, variance=None, iou_fn=None):
    """iou3d_nms_utils.py:227-273.  ``boxes`` (N, 7) and ``scores`` (N,) are numpy arrays that are
    updated in place, as in the reference.  Returns ``(scores, boxes)``."""
    if iou_fn is None:
        from .iou3d_nms_utils import boxes_bev_iou_cpu as iou_fn
    undone = scores >= score_threshold
    ious_all = iou_fn(boxes, boxes)                     # one N x N matrix from the ORIGINAL boxes
    two_pi = np.pi * 2
    while undone.sum() > 0:
        cand = undone.nonzero()[0]
        idx = cand[scores[cand].argmax()]
        if score_threshold <= 0 and scores[idx] <= 0:
            break                                        # only suppressed boxes are left (see module docstring)
        ious = ious_all[undone, idx]
        if variance is not None:
            top = boxes[idx]
            sel = ious > iou_threshold
            klbox = boxes[undone][sel]
            wrap = np.abs(klbox[:, 6] - top[6]) >= np.pi * 3 / 2
            klbox[wrap, 6] += two_pi if top[6] > 0 else -two_pi
            kliou = ious[sel]
            klvar = variance[undone, :7][sel]
            w = np.exp(-1 * (1 - kliou) ** 2 / _STD_IOU_SIGMA).reshape(-1, 1)
            w = w / klvar
            w[np.abs(klbox[:, 6] - top[6]) >= np.pi / 4, 6] = 0
            w = w / w.sum(0)
            boxes[idx, :7] = (w * klbox[:, :7]).sum(0)
        undone[idx] = False
        scores[undone] *= (ious_all[undone, idx] < iou_threshold)
        undone[scores < score_threshold] = False
    return scores, boxes


def new_nms_gpu(boxes, scores, iou_threshold, pre_maxsize=None, score_threshold=0, variance=None, **kwargs):
    """
    :param boxes: (N, 7) [x, y, z, dx, dy, dz, heading]
    :param scores: (N)
    :param thresh:
    :return: (keep indices sorted by descending new score [numpy], None, voted boxes [numpy])

    iou3d_nms_utils.py:200-224.  ``**kwargs`` swallows the NMS_CONFIG dict (model_nms_utils.py:40-43).
    """
    boxes = boxes.detach().cpu().numpy()
    scores = scores.detach().cpu().numpy()
    variance = variance.detach().cpu().numpy() if variance is not None else None
    boxes[:, 6] = _limit_period(boxes[:, 6], offset=0.5, period=np.pi * 2)
    new_scores, new_boxes = nms_func(boxes, scores, iou_threshold, score_threshold, variance=variance)
    keep = (new_scores > 0).nonzero()[0]
    keep = keep[new_scores[keep].argsort()[::-1]]
    return keep, None, new_boxes


def scale_by_iou(ious, soft_sigma, soft_mode="gaussian"):
    """iou3d_nms_utils.py:303-310."""
    if soft_mode == "linear":
        scale = ious.new_ones(ious.size())
        scale[ious >= soft_sigma] = 1 - ious[ious >= soft_sigma]
    else:
        scale = torch.exp(-ious ** 2 / soft_sigma)
    return scale


def softnms(boxes, scores, iou_threshold, soft_sigma, score_threshold, soft_mode="gaussian", variance=None):
    """iou3d_nms_utils.py:312-356: one boxes_iou_bev launch per iteration against the CURRENT boxes."""
    from .iou3d_nms_utils import boxes_iou_bev
    assert soft_mode in ["linear", "gaussian"]
    undone = scores >= score_threshold
    while undone.sum() > 1:
        idx = scores[undone].argmax()
        idx = undone.nonzero(as_tuple=False)[idx].item()
        top_box = boxes[idx:idx + 1]
        undone[idx] = False
        cur = boxes[undone]
        ious = boxes_iou_bev(cur, top_box).flatten()
        if variance is not None:
            sel = ious > iou_threshold
            klbox = torch.cat((cur[sel], top_box), 0)
            klvar = torch.cat((variance[undone, :6][sel], variance[idx:idx + 1, :6]), 0)
            w = torch.exp(-1 * torch.pow((1 - ious[sel]), 2) / _STD_IOU_SIGMA)
            w = torch.cat((w, torch.ones(1, device=w.device, dtype=w.dtype)), 0).unsqueeze(1)
            w = w / klvar
            w = w / w.sum(0)
            boxes[idx, :6] = (w * klbox[:, :6]).sum(0)
        scores[undone] *= scale_by_iou(ious, soft_sigma, soft_mode).flatten()
        undone[scores < score_threshold] = False
    return scores, boxes


def softnms_gpu(boxes, scores, iou_threshold, score_threshold=0.1, soft_mode='gaussian', variance=None, soft_sigma=0.3, **kwargs):
    """iou3d_nms_utils.py:292-301."""
    assert soft_mode in ["linear", "gaussian"]
    assert boxes.shape[-1] == 7
    new_scores, new_boxes = softnms(boxes, scores, iou_threshold, soft_sigma, score_threshold, soft_mode, variance=variance)
    keep = (new_scores > score_threshold).nonzero(as_tuple=False).view(-1)
    keep = keep[new_scores[keep].argsort(descending=True)]
    return keep, None, new_boxes
