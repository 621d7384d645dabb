"""Drop-in for the row-aligned IoU of ``pcdet.ops.iou3d.iou3d_utils``.

GLENet's IoU-aware heads call ``boxes_aligned_iou3d_gpu(pred_boxes[pos], gt_boxes[pos])`` once per training step
(``pcdet/models/dense_heads/anchor_head_kl_label.py:428``, ``anchor_head_iou.py:209``).  The reference builds the result
from ``boxes3d_to_bev_torch`` (``iou3d_utils.py:79-106``), the native ``boxes_aligned_overlap_bev_gpu``
(``pcdet/ops/iou3d/src/iou3d.cpp:55-73``, one 16-thread block per 16 pairs) and ~25 torch elementwise kernels
(``iou3d_utils.py:332-387``); here it is one kernel launch with the same per-step float32 rounding.

This op is NOT the arithmetic of ``pcdet.ops.iou3d_nms``: boxes are ``[x1, y1, x2, y2, ry]`` rectangles rotated
clockwise about their centre, and ``check_in_box2d`` uses a 1e-5 margin on the box edges
(``iou3d_kernel.cu:50-66,122-126``) -- reproduced from the reference kernel's SASS, see ``csrc/geom.cuh``.
"""
from __future__ import annotations

import torch

from . import _lib
from .iou3d_nms_utils import _check_cuda_f32, _device_for_host_call, _stream

__all__ = ["boxes_aligned_iou3d_gpu", "boxes_aligned_overlap_bev_gpu", "boxes_aligned_overlap_bev_cpu", "boxes3d_to_bev_torch"]


def boxes3d_to_bev_torch(boxes3d, box_mode='wlh', rect=False):
    """(N, 7) ``[x, y, z, d3, d4, d5, ry]`` (or (N, 5) ``[x, y, d2, d3, ry]``) -> (N, 5) ``[x1, y1, x2, y2, ry]``.

    Same result as the reference helper (``pcdet/ops/iou3d/iou3d_utils.py:79-106``): ``box_mode`` names which of the three
    extent columns is 'w' (the extent along the first BEV axis) and 'l' (along the second); ``rect=True`` selects the
    camera convention (BEV plane = columns 0 and 2, 'l' along the first axis).  One fused ``stack`` instead of five
    column assignments; the halves are computed as ``extent / 2`` and added / subtracted once each, as there.
    :func:`boxes_aligned_iou3d_gpu` does NOT call this -- the kernel fuses the conversion -- it is kept for callers
    that want the BEV rows (``boxes_aligned_overlap_bev_gpu``)."""
    width = boxes3d.shape[-1]
    if width not in (5, 7):
        raise NotImplementedError
    first_extent = 2 if width == 5 else 3
    half = {k: boxes3d[:, first_extent + box_mode.index(k)] / 2. for k in 'wl'}
    u, v = boxes3d[:, 0], boxes3d[:, 2 if rect else 1]
    hu, hv = (half['l'], half['w']) if rect else (half['w'], half['l'])
    return torch.stack((u - hu, v - hv, u + hu, v + hv, boxes3d[:, -1]), dim=1)


def boxes_aligned_overlap_bev_gpu(boxes_a_bev, boxes_b_bev):
    """The native call of the reference (iou3d.cpp:55-73): (N, 5) x (N, 5) ``[x1, y1, x2, y2, ry]`` -> (N, 1) overlap areas."""
    _check_cuda_f32(boxes_a_bev, "boxes_a")
    _check_cuda_f32(boxes_b_bev, "boxes_b")
    assert boxes_a_bev.shape == boxes_b_bev.shape and boxes_a_bev.shape[1] == 5
    a, b = boxes_a_bev.contiguous(), boxes_b_bev.contiguous()
    out = torch.empty((a.shape[0], 1), dtype=torch.float32, device=a.device)
    if a.shape[0]:
        lib = _lib.load()
        with torch.cuda.device(a.device):
            rc = lib.glenet_iou3d_v1_aligned_overlap_bev_gpu(a.data_ptr(), b.data_ptr(), a.shape[0], out.data_ptr(), _stream(a.device))
        _lib.check(rc, "glenet_iou3d_v1_aligned_overlap_bev_gpu")
    return out


def boxes_aligned_iou3d_gpu(boxes_a, boxes_b, box_mode='wlh', rect=False, need_bev=False):
    """
    Input (torch):
        boxes_a: (N, 7) [x, y, z, w, l, h, ry], torch tensor with type float32.
        boxes_b: (N, 7) [x, y, z, w, l, h, ry], torch tensor with type float32.
        rect: True/False means boxes in camera/velodyne coord system.
        Notice: (x, y, z) are real center.
    Output:
        iou_3d: (N, 1)   [and iou_bev: (N, 1) with need_bev]

    Reference: iou3d_utils.py:332-387.  ``rect=True`` raises NotImplementedError there as well (:356-357).
    """
    assert boxes_a.shape[0] == boxes_b.shape[0]
    w_index, l_index, h_index = box_mode.index('w') + 3, box_mode.index('l') + 3, box_mode.index('h') + 3
    if rect:
        raise NotImplementedError
    _check_cuda_f32(boxes_a, "boxes_a")
    _check_cuda_f32(boxes_b, "boxes_b")
    assert boxes_a.dim() == 2 and boxes_a.shape[1] == 7 and boxes_b.shape[1] == 7
    a, b = boxes_a.contiguous(), boxes_b.contiguous()
    n = a.shape[0]
    iou3d = torch.empty((n, 1), dtype=torch.float32, device=a.device)
    iou_bev = torch.empty((n, 1), dtype=torch.float32, device=a.device) if need_bev else None
    if n:
        lib = _lib.load()
        with torch.cuda.device(a.device):
            rc = lib.glenet_iou3d_v1_boxes_aligned_gpu(a.data_ptr(), b.data_ptr(), n, w_index, l_index, h_index, iou3d.data_ptr(),
                                                       iou_bev.data_ptr() if need_bev else None, None, _stream(a.device))
        _lib.check(rc, "glenet_iou3d_v1_boxes_aligned_gpu")
    if need_bev:
        return iou3d, iou_bev
    return iou3d


def boxes_aligned_overlap_bev_cpu(boxes_a_bev, boxes_b_bev):
    """Row-aligned counterpart of the op's CPU function ``boxes_overlap_bev_cpu`` (``pcdet/ops/iou3d/src/iou3d_cpu.cpp:258-281``):
    CPU tensors (N, 5) ``[x1, y1, x2, y2, ry]`` in, (N, 1) CPU tensor out, element i = ``box_overlap(a[i], b[i])`` of
    ``iou3d_cpu.cpp:126-247`` bit for bit.  Executes on the GPU in that file's CPU dialect (every operation rounded
    separately; cos/sin of the angles evaluated by the host's libm, as for ``iou3d_nms_utils.boxes_bev_iou_cpu``)."""
    assert not (boxes_a_bev.is_cuda or boxes_b_bev.is_cuda), 'Only support CPU tensors'
    assert boxes_a_bev.shape == boxes_b_bev.shape and boxes_a_bev.shape[1] == 5
    if boxes_a_bev.dtype != torch.float32 or boxes_b_bev.dtype != torch.float32:
        raise RuntimeError("boxes must be float32")
    a, b = boxes_a_bev.contiguous(), boxes_b_bev.contiguous()
    n = a.shape[0]
    ans = a.new_zeros((n, 1))
    if n:
        dev = _device_for_host_call("boxes_aligned_overlap_bev_cpu")
        lib = _lib.load()
        # one pinned host buffer [boxes_a | boxes_b | pad | trig_a | trig_b] -> one H2D copy; trig tables 16-byte aligned
        o_b, o_ta = n * 5, (2 * n * 5 + 3) // 4 * 4
        o_tb = o_ta + 4 * n
        host = torch.empty((o_tb + 4 * n,), dtype=torch.float32).pin_memory()
        host[:o_b].copy_(a.view(-1))
        host[o_b:o_b + n * 5].copy_(b.view(-1))
        base = host.data_ptr()
        lib.glenet_host_trig4_strided(a.data_ptr() + 16, 5, n, base + 4 * o_ta)
        lib.glenet_host_trig4_strided(b.data_ptr() + 16, 5, n, base + 4 * o_tb)
        d = host.to(dev, non_blocking=True)
        out = torch.empty((n, 1), dtype=torch.float32, device=dev)
        p = d.data_ptr()
        with torch.cuda.device(dev):
            rc = lib.glenet_iou3d_v1_aligned_overlap_bev_cpu_dialect(p, p + 4 * o_ta, p + 4 * o_b, p + 4 * o_tb, n, out.data_ptr(), _stream(dev))
        _lib.check(rc, "glenet_iou3d_v1_aligned_overlap_bev_cpu_dialect")
        ans.copy_(out)   # D2H, synchronising
    return ans
