"""Drop-in for the points-in-boxes functions of ``pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils``
(``pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:9-41``), backed by ``libglenet_geom.so``.

``RoIAwarePool3d`` (the PartA2 voxel pooling layer, same file :44-107) is out of scope of
this hot path and is intentionally not provided.
"""
from __future__ import annotations

import torch

from . import _lib
from .iou3d_nms_utils import _device_for_host_call, _stream, check_numpy_to_torch

__all__ = ["points_in_boxes_cpu", "points_in_boxes_cpu_lists", "points_in_boxes_gpu"]


def points_in_boxes_cpu(points, boxes):
    """
    Args:
        points: (num_points, 3)
        boxes: [x, y, z, dx, dy, dz, heading], (x, y, z) is the box center, each box DO NOT overlaps
    Returns:
        point_indices: (N, num_points)

    Reference: roiaware_pool3d_utils.py:9-25 -> points_in_boxes_cpu (roiaware_pool3d.cpp:143-168),
    a single-threaded N x num_points host loop with MARGIN = 1e-2.  Same signature and result
    (int32 0/1 mask, numpy iff ``boxes`` is numpy), but evaluated on the GPU in the CPU dialect:
    host-libm cos/sin per box, products rounded separately, FP64 comparisons.
    """
    assert boxes.shape[1] == 7
    assert points.shape[1] == 3
    points, is_numpy = check_numpy_to_torch(points)
    boxes, is_numpy = check_numpy_to_torch(boxes)
    n, m = boxes.shape[0], points.shape[0]
    point_indices = points.new_zeros((n, m), dtype=torch.int)
    if n and m:
        point_indices.copy_(_cpu_dialect_mask_on_device(points, boxes))   # D2H, synchronising
    return point_indices.numpy() if is_numpy else point_indices


def _cpu_dialect_mask_on_device(points: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
    """(N, M) int32 0/1 mask of points_in_boxes_cpu, left on the GPU."""
    n, m = boxes.shape[0], points.shape[0]
    b = boxes.float().contiguous()
    p = points.float().contiguous()
    if b.is_cuda or p.is_cuda:
        raise RuntimeError("points_in_boxes_cpu expects CPU tensors / numpy arrays")
    dev = _device_for_host_call("points_in_boxes_cpu")
    lib = _lib.load()
    host = torch.empty(n * 7 + n * 2 + m * 3, dtype=torch.float32).pin_memory()
    o_t, o_p = n * 7, n * 9
    host[:o_t].copy_(b.view(-1))
    lib.glenet_host_trig2(b.data_ptr(), n, host.data_ptr() + 4 * o_t)
    host[o_p:].copy_(p.view(-1))
    d = host.to(dev, non_blocking=True)
    out = torch.empty((n, m), dtype=torch.int32, device=dev)
    base = d.data_ptr()
    with torch.cuda.device(dev):
        rc = lib.glenet_points_in_boxes_cpu_dialect(base, base + 4 * o_t, n, base + 4 * o_p, m, out.data_ptr(), _stream(dev))
    _lib.check(rc, "glenet_points_in_boxes_cpu_dialect")
    return out


def points_in_boxes_cpu_lists(points, boxes):
    """Per-box index lists instead of the (N, num_points) mask of :func:`points_in_boxes_cpu`.

    Returns ``(offsets, indices)``: ``indices[offsets[i]:offsets[i + 1]]`` are, in ascending order, the points with
    ``points_in_boxes_cpu(points, boxes)[i] > 0`` -- i.e. ``np.nonzero(point_indices[i])[0]``, the selection the
    GT-database builders make per object (``kitti_dataset.py:236-259``: ``gt_points = points[point_indices[i] > 0]``,
    ``waymo_dataset.py:342-395``).  Same predicate and dialect as points_in_boxes_cpu (MARGIN 1e-2, a point may belong to
    several boxes); the mask is compacted on the GPU, so only the few thousand indices of the objects' points cross
    PCIe instead of N x num_points int32.  numpy in -> numpy out, like the mask function.  Additive API."""
    assert boxes.shape[1] == 7
    assert points.shape[1] == 3
    points, is_numpy = check_numpy_to_torch(points)
    boxes, is_numpy = check_numpy_to_torch(boxes)
    n, m = boxes.shape[0], points.shape[0]
    offsets = torch.zeros((n + 1,), dtype=torch.int64)
    indices = torch.zeros((0,), dtype=torch.int64)
    if n and m:
        mask = _cpu_dialect_mask_on_device(points, boxes)
        nz = mask.nonzero()                                   # row-major: boxes ascending, points ascending within a box
        counts = torch.bincount(nz[:, 0], minlength=n)
        offsets[1:] = torch.cumsum(counts, 0).cpu()
        indices = nz[:, 1].contiguous().cpu()
    return (offsets.numpy(), indices.numpy()) if is_numpy else (offsets, indices)


def points_in_boxes_gpu(points, boxes):
    """
    :param points: (B, M, 3)
    :param boxes: (B, T, 7), num_valid_boxes <= T
    :return box_idxs_of_pts: (B, M), default background = -1

    Reference: roiaware_pool3d_utils.py:28-41 -> points_in_boxes_gpu (roiaware_pool3d.cpp:98-118).
    """
    assert boxes.shape[0] == points.shape[0]
    assert boxes.shape[2] == 7 and points.shape[2] == 3
    if not (points.is_cuda and boxes.is_cuda):
        raise RuntimeError("points_in_boxes_gpu expects CUDA tensors")
    if points.dtype != torch.float32 or boxes.dtype != torch.float32:
        raise RuntimeError("points_in_boxes_gpu expects float32 tensors")   # reference: .data<float>() throws
    if points.device != boxes.device:
        raise RuntimeError("points and boxes must be on the same device")
    batch_size, num_points, _ = points.shape
    num_boxes = boxes.shape[1]
    dev = points.device
    box_idxs_of_pts = torch.empty((batch_size, num_points), dtype=torch.int32, device=dev)
    if batch_size and num_points:
        lib = _lib.load()
        b = boxes.contiguous()
        p = points.contiguous()
        ws_bytes = lib.glenet_points_in_boxes_workspace_bytes(batch_size, num_boxes)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.glenet_points_in_boxes_gpu(b.data_ptr(), p.data_ptr(), batch_size, num_boxes, num_points,
                                                box_idxs_of_pts.data_ptr(), ws.data_ptr(), ws_bytes, _stream(dev))
        _lib.check(rc, "glenet_points_in_boxes_gpu")
    return box_idxs_of_pts
