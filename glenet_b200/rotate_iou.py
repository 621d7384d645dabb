"""Drop-in for ``pcdet.datasets.kitti.kitti_object_eval_python.rotate_iou`` (the KITTI evaluator's rotated BEV IoU).

``rotate_iou_gpu_eval`` keeps the reference's signature and behaviour (``rotate_iou.py:263-330``: numpy in, numpy out in the
input dtype, float32 arithmetic, ``criterion`` -1 / 0 / 1 / other, ``device_id``) but runs ``csrc/rotate_iou.cu`` through the
C ABI (``glenet_rotate_iou_eval_gpu``) instead of a numba-CUDA kernel.  ``bev_box_overlap`` / ``d3_box_overlap`` are the two
callers in ``kitti_object_eval_python/eval.py:115-151``.

Additive: :func:`rotate_iou_gpu_eval_blocks` computes only the per-frame diagonal blocks that ``calculate_iou_partly``
(``eval.py:344-400``) slices out of the (sum of GT) x (sum of detections) matrix of one evaluation part.
There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

__all__ = ["rotate_iou_gpu_eval", "rotate_iou_gpu_eval_blocks", "bev_box_overlap", "d3_box_overlap"]


def _device(device_id: int) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("glenet_b200.rotate_iou needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", int(device_id))


def rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1, device_id=0):
    """rotated box iou running in gpu.

    Args:
        boxes (float array: [N, 5]): rbboxes. format: centers, dims, angles(clockwise when positive)
        query_boxes (float array: [K, 5])
        criterion: -1 IoU, 0 intersection / area(query box), 1 intersection / area(box), else the intersection area
        device_id (int, optional): Defaults to 0.
    Returns:
        (N, K) array of ``boxes.dtype``

    Reference: rotate_iou.py:263-330 (``iou[n, k] = devRotateIoUEval(query_boxes[k], boxes[n], criterion)``).
    """
    box_dtype = boxes.dtype
    boxes = np.ascontiguousarray(boxes.astype(np.float32))
    query_boxes = np.ascontiguousarray(query_boxes.astype(np.float32))
    n, k = boxes.shape[0], query_boxes.shape[0]
    iou = np.zeros((n, k), dtype=np.float32)
    if n == 0 or k == 0:
        return iou
    assert boxes.ndim == 2 and boxes.shape[1] == 5 and query_boxes.ndim == 2 and query_boxes.shape[1] == 5
    dev = _device(device_id)
    lib = _lib.load()
    host = torch.empty((n + k) * 5, dtype=torch.float32).pin_memory()
    host[: n * 5].copy_(torch.from_numpy(boxes).view(-1))
    host[n * 5:].copy_(torch.from_numpy(query_boxes).view(-1))
    with torch.cuda.device(dev):
        d = host.to(dev, non_blocking=True)
        out = torch.empty((n, k), dtype=torch.float32, device=dev)
        rc = lib.glenet_rotate_iou_eval_gpu(d.data_ptr(), n, d.data_ptr() + 4 * n * 5, k, int(criterion), out.data_ptr(),
                                            torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "glenet_rotate_iou_eval_gpu")
        iou = out.cpu().numpy()          # D2H, synchronising (the reference's stream.auto_synchronize())
    return iou.astype(box_dtype)


def rotate_iou_gpu_eval_blocks(boxes, query_boxes, box_counts, query_counts, criterion=-1, device_id=0):
    """Block-diagonal form: group g pairs ``boxes[sum(box_counts[:g]):][:box_counts[g]]`` with the matching slice of
    ``query_boxes``.  Returns the list of (box_counts[g], query_counts[g]) float32 arrays -- exactly the blocks
    ``calculate_iou_partly`` (eval.py:383-397) cuts out of the dense matrix, without computing the off-diagonal part.
    One launch for all groups (``glenet_rotate_iou_eval_blocks_gpu``)."""
    boxes = np.ascontiguousarray(np.asarray(boxes).astype(np.float32)).reshape(-1, 5)
    query_boxes = np.ascontiguousarray(np.asarray(query_boxes).astype(np.float32)).reshape(-1, 5)
    bc = np.asarray(box_counts, dtype=np.int64)
    qc = np.asarray(query_counts, dtype=np.int64)
    assert bc.ndim == 1 and bc.shape == qc.shape and (bc >= 0).all() and (qc >= 0).all()
    assert bc.sum() == boxes.shape[0] and qc.sum() == query_boxes.shape[0]
    groups = bc.shape[0]
    boff = np.zeros(groups + 1, dtype=np.int32); boff[1:] = np.cumsum(bc)
    qoff = np.zeros(groups + 1, dtype=np.int32); qoff[1:] = np.cumsum(qc)
    ooff = np.zeros(groups + 1, dtype=np.int64); ooff[1:] = np.cumsum(bc * qc)
    total = int(ooff[-1])
    flat = np.zeros((total,), dtype=np.float32)
    if total:
        dev = _device(device_id)
        lib = _lib.load()
        with torch.cuda.device(dev):
            d_b = torch.from_numpy(boxes).to(dev)
            d_q = torch.from_numpy(query_boxes).to(dev)
            d_bo, d_qo, d_oo = torch.from_numpy(boff).to(dev), torch.from_numpy(qoff).to(dev), torch.from_numpy(ooff).to(dev)
            out = torch.empty((total,), dtype=torch.float32, device=dev)
            rc = lib.glenet_rotate_iou_eval_blocks_gpu(d_b.data_ptr(), d_bo.data_ptr(), d_q.data_ptr(), d_qo.data_ptr(), d_oo.data_ptr(), groups,
                                                       int(bc.max()), int(qc.max()), int(criterion), out.data_ptr(),
                                                       torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc, "glenet_rotate_iou_eval_blocks_gpu")
            flat = out.cpu().numpy()
    return [flat[ooff[g]:ooff[g + 1]].reshape(int(bc[g]), int(qc[g])) for g in range(groups)]


def bev_box_overlap(boxes, qboxes, criterion=-1):
    """kitti_object_eval_python/eval.py:115-117."""
    return rotate_iou_gpu_eval(boxes, qboxes, criterion)


def d3_box_overlap(boxes, qboxes, criterion=-1):
    """kitti_object_eval_python/eval.py:120-151: BEV intersection area (criterion 2) times the overlap of the camera-frame
    height intervals, over the volume union.  The height step follows ``d3_box_overlap_kernel`` (a numba CPU loop in the
    reference) elementwise in the arrays' own dtype."""
    rinc = rotate_iou_gpu_eval(boxes[:, [0, 2, 3, 5, 6]], qboxes[:, [0, 2, 3, 5, 6]], 2)
    if rinc.size == 0:
        return rinc
    b, q = boxes, qboxes
    iw = np.minimum(b[:, None, 1], q[None, :, 1]) - np.maximum(b[:, None, 1] - b[:, None, 4], q[None, :, 1] - q[None, :, 4])
    area1 = (b[:, 3] * b[:, 4] * b[:, 5])[:, None]
    area2 = (q[:, 3] * q[:, 4] * q[:, 5])[None, :]
    inc = iw * rinc
    if criterion == -1:
        ua = area1 + area2 - inc
    elif criterion == 0:
        ua = np.broadcast_to(area1, inc.shape)
    elif criterion == 1:
        ua = np.broadcast_to(area2, inc.shape)
    else:
        ua = inc
    pos = (rinc > 0) & (iw > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        val = inc / ua
    out = np.where(rinc > 0, np.where(pos, val, 0.0), rinc)
    return out.astype(rinc.dtype)
