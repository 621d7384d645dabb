"""Loader for the compiled reference (``oracle/_ref``) -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``glenet_b200``) never does.

``oracle/_ref/*.so`` are the reference's own pybind11 modules, built unmodified
by ``oracle/build_ref.py``.  The thin Python functions below restate what the
reference's wrapper modules do around those native calls so that the oracle is
called with exactly the reference's argument preparation:

* ``pcdet/ops/iou3d_nms/iou3d_nms_utils.py:52-121,182-197,276-290``
* ``pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:9-41``
* ``pcdet/utils/common_utils.py:15-18`` (``check_numpy_to_torch``)

(``tests/test_oracle_ref.py`` checks these restatements against the reference's
wrapper files imported verbatim, when ``/root/reference`` is present.)
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_mods = {}


def available() -> bool:
    return all(os.path.isfile(os.path.join(REF_DIR, n + ".so")) for n in ("iou3d_nms_cuda", "roiaware_pool3d_cuda"))


def _load(name: str):
    if name not in _mods:
        path = os.path.join(REF_DIR, name + ".so")
        if not os.path.isfile(path):
            raise FileNotFoundError(f"{path} missing: run `python oracle/build_ref.py` where /root/reference exists")
        loader = importlib.machinery.ExtensionFileLoader(name, path)
        spec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _mods[name] = mod
    return _mods[name]


def iou3d_nms_cuda():
    return _load("iou3d_nms_cuda")


def roiaware_pool3d_cuda():
    return _load("roiaware_pool3d_cuda")


def iou3d_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "iou3d_cuda.so"))


def iou3d_cuda():
    """pcdet/ops/iou3d (the [x1, y1, x2, y2, ry] variant used by boxes_aligned_iou3d_gpu)."""
    return _load("iou3d_cuda")


def rotate_iou_numba():
    """The reference's numba-CUDA module (kitti_object_eval_python/rotate_iou.py), staged unmodified by build_ref.py.
    Importing it JIT-compiles the kernel for the current device: GPU box only.  None when unavailable."""
    path = os.path.join(REF_DIR, "rotate_iou_numba.py")
    if "rotate_iou_numba" not in _mods:
        mod = None
        if os.path.isfile(path):
            try:
                spec = importlib.util.spec_from_file_location("rotate_iou_numba", path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
            except Exception:   # no GPU / numba without CUDA support
                mod = None
        _mods["rotate_iou_numba"] = mod
    return _mods["rotate_iou_numba"]


def _np2t(x):
    # common_utils.py:15-18
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).float(), True
    return x, False


# ---------------------------------------------------------------- CPU dialect
def boxes_bev_iou_cpu(boxes_a, boxes_b):
    """iou3d_nms_utils.py:52-68."""
    boxes_a, is_numpy = _np2t(boxes_a)
    boxes_b, is_numpy = _np2t(boxes_b)
    assert not (boxes_a.is_cuda or boxes_b.is_cuda)
    assert boxes_a.shape[1] == 7 and boxes_b.shape[1] == 7
    ans = boxes_a.new_zeros(torch.Size((boxes_a.shape[0], boxes_b.shape[0])))
    iou3d_nms_cuda().boxes_iou_bev_cpu(boxes_a.contiguous(), boxes_b.contiguous(), ans)
    return ans.numpy() if is_numpy else ans


def points_in_boxes_cpu(points, boxes):
    """roiaware_pool3d_utils.py:9-25."""
    assert boxes.shape[1] == 7 and points.shape[1] == 3
    points, is_numpy = _np2t(points)
    boxes, is_numpy = _np2t(boxes)
    out = points.new_zeros((boxes.shape[0], points.shape[0]), dtype=torch.int)
    roiaware_pool3d_cuda().points_in_boxes_cpu(boxes.float().contiguous(), points.float().contiguous(), out)
    return out.numpy() if is_numpy else out


# ---------------------------------------------------------------- GPU dialect (needs a GPU)
def boxes_iou_bev(boxes_a, boxes_b):
    """iou3d_nms_utils.py:71-85."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    ans = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    if ans.numel():
        iou3d_nms_cuda().boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans)
    return ans


def boxes_overlap_bev(boxes_a, boxes_b):
    ans = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    if ans.numel():
        iou3d_nms_cuda().boxes_overlap_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans)
    return ans


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """iou3d_nms_utils.py:88-121 (each elementwise step separately rounded, as torch does)."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    a_hmax = (boxes_a[:, 2] + boxes_a[:, 5] / 2).view(-1, 1)
    a_hmin = (boxes_a[:, 2] - boxes_a[:, 5] / 2).view(-1, 1)
    b_hmax = (boxes_b[:, 2] + boxes_b[:, 5] / 2).view(1, -1)
    b_hmin = (boxes_b[:, 2] - boxes_b[:, 5] / 2).view(1, -1)
    overlaps_bev = boxes_overlap_bev(boxes_a, boxes_b)
    max_of_min = torch.max(a_hmin, b_hmin)
    min_of_max = torch.min(a_hmax, b_hmax)
    overlaps_h = torch.clamp(min_of_max - max_of_min, min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)


def _nms(fn, boxes, scores, thresh, pre_maxsize=None):
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    keep = torch.zeros(boxes.size(0), dtype=torch.int64)
    num_out = fn(boxes, keep, thresh) if boxes.size(0) else 0
    return order[keep[:num_out].to(boxes.device)].contiguous(), None


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """iou3d_nms_utils.py:182-197."""
    return _nms(iou3d_nms_cuda().nms_gpu, boxes, scores, thresh, pre_maxsize)


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    """iou3d_nms_utils.py:276-290."""
    return _nms(iou3d_nms_cuda().nms_normal_gpu, boxes, scores, thresh)


def points_in_boxes_gpu(points, boxes):
    """roiaware_pool3d_utils.py:28-41."""
    assert boxes.shape[0] == points.shape[0]
    assert boxes.shape[2] == 7 and points.shape[2] == 3
    b, m, _ = points.shape
    out = points.new_zeros((b, m), dtype=torch.int).fill_(-1)
    if out.numel() and boxes.shape[1]:
        roiaware_pool3d_cuda().points_in_boxes_gpu(boxes.contiguous(), points.contiguous(), out)
    return out


# ---------------------------------------------------------------- pcdet/ops/iou3d (row-aligned IoU of the IoU-aware heads)
def boxes3d_to_bev_torch(boxes3d, box_mode='wlh', rect=False):
    """pcdet/ops/iou3d/iou3d_utils.py:79-106."""
    boxes_bev = boxes3d.new(torch.Size((boxes3d.shape[0], 5)))
    w_index, l_index = box_mode.index('w') + 3, box_mode.index('l') + 3
    half_w, half_l = boxes3d[:, w_index] / 2., boxes3d[:, l_index] / 2.
    assert not rect
    cu, cv = boxes3d[:, 0], boxes3d[:, 1]
    boxes_bev[:, 0], boxes_bev[:, 1] = cu - half_w, cv - half_l
    boxes_bev[:, 2], boxes_bev[:, 3] = cu + half_w, cv + half_l
    boxes_bev[:, 4] = boxes3d[:, -1]
    return boxes_bev


def boxes_aligned_iou3d_gpu(boxes_a, boxes_b, box_mode='wlh', rect=False, need_bev=False):
    """pcdet/ops/iou3d/iou3d_utils.py:332-387 (rect=False; rect=True raises there too)."""
    assert boxes_a.shape[0] == boxes_b.shape[0]
    w_index, l_index, h_index = box_mode.index('w') + 3, box_mode.index('l') + 3, box_mode.index('h') + 3
    boxes_a_bev = boxes3d_to_bev_torch(boxes_a, box_mode, rect)
    boxes_b_bev = boxes3d_to_bev_torch(boxes_b, box_mode, rect)
    overlaps_bev = torch.zeros((boxes_a.shape[0], 1), dtype=torch.float32, device=boxes_a.device)
    if boxes_a.shape[0]:
        iou3d_cuda().boxes_aligned_overlap_bev_gpu(boxes_a_bev.contiguous(), boxes_b_bev.contiguous(), overlaps_bev)
    area_a = (boxes_a[:, w_index] * boxes_a[:, l_index]).view(-1, 1)
    area_b = (boxes_b[:, w_index] * boxes_b[:, l_index]).view(-1, 1)
    iou_bev = overlaps_bev / torch.clamp(area_a + area_b - overlaps_bev, min=1e-7)
    if rect:
        raise NotImplementedError
    half_h_a = boxes_a[:, h_index] / 2.0
    half_h_b = boxes_b[:, h_index] / 2.0
    a_hmin = (boxes_a[:, 2] - half_h_a).view(-1, 1)
    a_hmax = (boxes_a[:, 2] + half_h_a).view(-1, 1)
    b_hmin = (boxes_b[:, 2] - half_h_b).view(-1, 1)
    b_hmax = (boxes_b[:, 2] + half_h_b).view(-1, 1)
    max_of_min = torch.max(a_hmin, b_hmin)
    min_of_max = torch.min(a_hmax, b_hmax)
    overlaps_h = torch.clamp(min_of_max - max_of_min, min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(-1, 1)
    iou3d = overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-7)
    if need_bev:
        return iou3d, iou_bev
    return iou3d


def iou3d_v1_overlap_bev_cpu(boxes_a_bev, boxes_b_bev):
    """boxes_overlap_bev_cpu of pcdet/ops/iou3d/src/iou3d_cpu.cpp:258-281: (N, 5) x (M, 5) CPU tensors -> (N, M)."""
    ans = torch.zeros((boxes_a_bev.shape[0], boxes_b_bev.shape[0]), dtype=torch.float32)
    iou3d_cuda().boxes_overlap_bev_cpu(boxes_a_bev.contiguous(), boxes_b_bev.contiguous(), ans)
    return ans


# ---------------------------------------------------------------- GLENet's variance-voting NMS / soft-NMS (Python loops of the reference)
def nms_func_reference(boxes, scores, iou_threshold, score_threshold=0, variance=None, ious_all=None):
    """iou3d_nms_utils.py:227-273 restated on numpy float32 arrays (updated in place, like the reference).  ``ious_all`` is the
    (N, N) matrix the reference gets from boxes_bev_iou_cpu(boxes, boxes); pass it in to check a device loop against the same
    matrix, or leave it None to compute it with the compiled reference.  The loop stops once only zero scores are left
    (those iterations change nothing the caller reads)."""
    if ious_all is None:
        ious_all = boxes_bev_iou_cpu(boxes, boxes)
    undone_mask = scores >= score_threshold
    while undone_mask.sum() > 0:
        idx = scores[undone_mask].argmax()
        idx = undone_mask.nonzero()[0][idx]
        if score_threshold <= 0 and not scores[idx] > 0:
            break
        top_box = boxes[idx:idx + 1]
        _boxes = boxes[undone_mask]
        ious = ious_all[undone_mask, idx]
        if variance is not None:
            _variance = variance[undone_mask, :7]
            ioumask = ious > iou_threshold
            klbox = _boxes[ioumask]
            if top_box[:, 6] > 0:
                klbox[np.abs(klbox[:, 6] - top_box[:, 6]) >= np.pi * 3 / 2, 6] += np.pi * 2
            else:
                klbox[np.abs(klbox[:, 6] - top_box[:, 6]) >= np.pi * 3 / 2, 6] -= np.pi * 2
            kliou = ious[ioumask]
            klvar = _variance[ioumask]
            pi = (np.exp(-1 * (1 - kliou) ** 2 / 0.05)).reshape(-1, 1)
            pi = pi / klvar
            pi[np.abs(klbox[:, 6] - top_box[:, 6]) >= np.pi / 4, 6] = 0
            pi = pi / pi.sum(0)
            boxes[idx, :7] = (pi * klbox[:, :7]).sum(0)
        undone_mask[idx] = False
        scores[undone_mask] *= (ious_all[undone_mask, idx] < iou_threshold)
        undone_mask[scores < score_threshold] = False
    return scores, boxes


def softnms_reference(boxes, scores, iou_threshold, soft_sigma, score_threshold, soft_mode="gaussian", variance=None):
    """iou3d_nms_utils.py:312-356 restated (CUDA tensors, in place): one reference boxes_iou_bev launch per iteration."""
    undone_mask = scores >= score_threshold
    while undone_mask.sum() > 1:
        idx = scores[undone_mask].argmax()
        idx = undone_mask.nonzero(as_tuple=False)[idx].item()
        top_box = boxes[idx:idx + 1]
        undone_mask[idx] = False
        _boxes = boxes[undone_mask]
        ious = boxes_iou_bev(_boxes, top_box).flatten()
        if variance is not None:
            _variance = variance[undone_mask, :6]
            ioumask = ious > iou_threshold
            klbox = torch.cat((_boxes[ioumask], top_box), 0)
            kliou = ious[ioumask]
            klvar = torch.cat((_variance[ioumask], variance[idx:idx + 1, :6]), 0)
            pi = torch.exp(-1 * torch.pow((1 - kliou), 2) / 0.05)
            pi = torch.cat((pi, torch.ones(1, device=pi.device)), 0).unsqueeze(1)
            pi = pi / klvar
            pi = pi / pi.sum(0)
            boxes[idx, :6] = (pi * klbox[:, :6]).sum(0)
        if soft_mode == "linear":
            scales = ious.new_ones(ious.size())
            scales[ious >= soft_sigma] = 1 - ious[ious >= soft_sigma]
        else:
            scales = torch.exp(-ious ** 2 / soft_sigma)
        scores[undone_mask] *= scales.flatten()
        undone_mask[scores < score_threshold] = False
    return scores, boxes


# ---------------------------------------------------------------- GT-database crops (kitti_dataset.py:248-254, waymo_dataset.py:369-372)
def gt_crops_reference(points: np.ndarray, gt_boxes: np.ndarray, selection: np.ndarray, rule: str):
    """The reference's per-object selection + centring, statement for statement: ``selection`` is the (N, M) mask of
    points_in_boxes_cpu (rule "kitti": ``points[point_indices[i] > 0]``) or the (M,) index vector of points_in_boxes_gpu
    (rule "waymo": ``points[box_idxs_of_pts == i]``); then ``gt_points[:, :3] -= gt_boxes[i, :3]`` in place on the float32
    rows (numpy subtracts in the boxes' dtype and casts back).  Returns the list of per-object arrays."""
    out = []
    for i in range(gt_boxes.shape[0]):
        gt_points = points[selection[i] > 0] if rule == "kitti" else points[selection == i]
        gt_points[:, :3] -= gt_boxes[i, :3]
        out.append(gt_points)
    return out
