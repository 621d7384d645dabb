"""GT-database crops on the GPU: points-in-boxes + stream compaction + the per-object ``.bin`` files.

The reference builds its ground-truth database (the sampling pool of ``gt_sampling`` and the on-disk input of the CVAE,
``cvae_uncertainty/dataset.py:313``) frame by frame on the host:

* KITTI (``pcdet/datasets/kitti/kitti_dataset.py:236-259``): ``points_in_boxes_cpu`` -> (objects, points) mask on the host,
  then per object ``gt_points = points[mask[i] > 0]; gt_points[:, :3] -= gt_boxes[i, :3]; gt_points.tofile(f)``;
* Waymo (``pcdet/datasets/waymo/waymo_dataset.py:364-380``): ``points_in_boxes_gpu`` -> (points,) index vector copied to the
  host, then per object ``gt_points = points[idx == i]; gt_points[:, :3] -= gt_boxes[i, :3]``.

Here the predicate (same kernels and dialects as the drop-in ``points_in_boxes_cpu`` / ``points_in_boxes_gpu``), the
selection and the centring run on the device (``csrc/crop.cu``, ``glenet_gt_crop_gpu``); what comes back over PCIe is one
float32 buffer holding every object's rows back to back plus an offset vector.  Bytes written to disk are identical to the
reference's.  Additive API; no CPU fallback.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .iou3d_nms_utils import _device_for_host_call, _stream
from .roiaware_pool3d_utils import _cpu_dialect_mask_on_device, points_in_boxes_gpu

__all__ = ["crop_points_in_boxes", "crop_gt_objects", "write_gt_crops", "CROP_MASK", "CROP_INDEX"]

CROP_MASK, CROP_INDEX = 0, 1


def crop_points_in_boxes(selection: torch.Tensor, points: torch.Tensor, centres: torch.Tensor, mode: int,
                         capacity: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Device-side selection + centring (``glenet_gt_crop_gpu``).

    Args:
        selection: CROP_MASK: (objects, points) int32 mask; CROP_INDEX: (points,) int32 index vector -- CUDA tensors
        points: (points, C) float32 CUDA tensor, C >= 3, xyz first
        centres: (objects, 3) float64 CUDA tensor
    Returns ``(offsets, crops)``: (objects + 1,) int64 and (offsets[-1], C) float32 CUDA tensors; rows
    ``offsets[i]:offsets[i + 1]`` are object i's points (ascending point index) with xyz relative to ``centres[i]``.
    One host synchronisation (the read of the total) because the result has a data-dependent length."""
    assert selection.is_cuda and points.is_cuda and centres.is_cuda
    assert selection.dtype == torch.int32 and points.dtype == torch.float32 and centres.dtype == torch.float64
    n_pts, feats = points.shape
    n_obj = centres.shape[0]
    assert centres.shape == (n_obj, 3) and feats >= 3
    assert selection.shape == ((n_obj, n_pts) if mode == CROP_MASK else (n_pts,))
    dev = points.device
    lib = _lib.load()
    sel, pts, ctr = selection.contiguous(), points.contiguous(), centres.contiguous()
    offsets = torch.empty((n_obj + 1,), dtype=torch.int64, device=dev)
    ws_bytes = lib.glenet_gt_crop_workspace_bytes(n_obj, n_pts)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    # an index vector selects every point at most once; a mask may select a point for several (overlapping) boxes
    cap = int(capacity) if capacity is not None else n_pts
    while True:
        crops = torch.empty((max(cap, 1), feats), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.glenet_gt_crop_gpu(mode, sel.data_ptr(), pts.data_ptr(), n_pts, feats, ctr.data_ptr(), n_obj, cap,
                                        offsets.data_ptr(), crops.data_ptr(), ws.data_ptr(), ws_bytes, _stream(dev))
        _lib.check(rc, "glenet_gt_crop_gpu")
        total = int(offsets[-1].item())
        if total <= cap:
            return offsets, crops[:total]
        cap = total                                   # overlapping boxes selected more rows than points: retry once, exactly sized


def crop_gt_objects(points, gt_boxes, rule: str = "kitti"):
    """The object crops of one frame, as the reference's database builders compute them.

    Args:
        points: (M, C) array / CPU tensor, float32, xyz first (KITTI C = 4, Waymo C = 5 or 6)
        gt_boxes: (N, >= 7) array / CPU tensor [x, y, z, dx, dy, dz, heading, ...], float32 or float64
        rule: ``"kitti"`` -- membership by ``points_in_boxes_cpu`` (MARGIN 1e-2, a point may fall in several boxes;
              kitti_dataset.py:248-254); ``"waymo"`` -- by ``points_in_boxes_gpu`` (first containing box; waymo_dataset.py:364-371)
    Returns ``(offsets, crops)`` numpy: int64 (N + 1,) and float32 (offsets[-1], C); ``crops[offsets[i]:offsets[i + 1]]`` equals
    the reference's ``gt_points`` of object i bit for bit."""
    pts = torch.as_tensor(np.asarray(points) if not torch.is_tensor(points) else points)
    box = torch.as_tensor(np.asarray(gt_boxes) if not torch.is_tensor(gt_boxes) else gt_boxes)
    assert pts.dim() == 2 and pts.shape[1] >= 3 and box.dim() == 2 and box.shape[1] >= 7
    assert pts.dtype == torch.float32, "the reference's point clouds are float32 (np.fromfile(..., dtype=np.float32))"
    n_obj, n_pts, feats = box.shape[0], pts.shape[0], pts.shape[1]
    if n_obj == 0 or n_pts == 0:
        return np.zeros((n_obj + 1,), dtype=np.int64), np.zeros((0, feats), dtype=np.float32)
    dev = _device_for_host_call("gt_database.crop_gt_objects")
    d_pts = pts.contiguous().pin_memory().to(dev, non_blocking=True)
    d_ctr = box[:, 0:3].to(torch.float64).contiguous().to(dev)       # exact for float32 boxes; float64 boxes subtract in float64 like numpy
    xyz = pts[:, 0:3]
    boxes7 = box[:, 0:7]
    if rule == "kitti":
        sel = _cpu_dialect_mask_on_device(xyz, boxes7)                 # (N, M) int32, stays on the device
        mode = CROP_MASK
    elif rule == "waymo":
        sel = points_in_boxes_gpu(d_pts[:, 0:3].contiguous().unsqueeze(0), boxes7.float().contiguous().to(dev).unsqueeze(0)).squeeze(0)
        mode = CROP_INDEX
    else:
        raise ValueError(f"unknown rule {rule!r}")
    offsets, crops = crop_points_in_boxes(sel, d_pts, d_ctr, mode)
    return offsets.cpu().numpy(), crops.cpu().numpy()


def write_gt_crops(database_save_path, filenames: Sequence[str], offsets: np.ndarray, crops: np.ndarray,
                   write: Optional[Sequence[bool]] = None) -> List[int]:
    """Write object i's rows to ``database_save_path / filenames[i]`` exactly as ``gt_points.tofile(f)`` does
    (kitti_dataset.py:256-257, waymo_dataset.py:376-377; file names: '%s_%s_%d.bin' % (sample_idx, name, i) and
    '%s_%04d_%s_%d.bin' % (sequence_name, sample_idx, name, i)).  Returns ``num_points_in_gt`` per object."""
    assert len(filenames) == len(offsets) - 1
    counts = []
    for i, name in enumerate(filenames):
        rows = crops[offsets[i]:offsets[i + 1]]
        counts.append(int(rows.shape[0]))
        if write is None or write[i]:
            with open(os.path.join(str(database_save_path), name), "w") as f:
                np.ascontiguousarray(rows).tofile(f)
    return counts
