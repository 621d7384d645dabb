"""Drop-in for ``pcdet.ops.iou3d_nms.iou3d_nms_utils`` backed by ``libglenet_geom.so``.

Same function names, argument meaning, return types and assertion behaviour as
``pcdet/ops/iou3d_nms/iou3d_nms_utils.py`` (reference lines cited per function).  Every
function executes hand-written sm_100a kernels through the C ABI of
``include/glenet_geom.h``; there is no CPU or eager fallback.

Differences that are deliberate and invisible to the reference's callers:

* outputs are allocated on the *inputs'* device and work is enqueued on torch's current
  stream of that device (the reference always uses the current device and the legacy
  default stream, ``iou3d_nms_utils.py:81,106`` / ``iou3d_nms_kernel.cu:383``);
* ``nms_gpu`` keeps the suppression mask and the greedy sweep on the device; the only
  host synchronisation is the read of the keep count that the variable-length return
  value requires (the reference does cudaMalloc/cudaFree, a blocking D2H copy of the whole
  mask and an H2D copy of ``keep``, ``iou3d_nms.cpp:103-114`` / ``iou3d_nms_utils.py:195-197``);
* N == 0 returns empty results instead of printing a CUDA launch error.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

__all__ = [
    "boxes_bev_iou_cpu", "boxes_iou_bev", "boxes_iou3d_gpu", "boxes_overlap_bev",
    "boxes_iou_bev_aligned", "boxes_iou3d_aligned",
    "boxes_iou_bev_frames", "boxes_overlap_bev_frames", "boxes_iou3d_gpu_frames",
    "boxes_iou_frames_sparse", "iou_max_overlaps_frames",
    "nms_gpu", "nms_normal_gpu", "nms_gpu_batch", "nms_normal_gpu_batch",
    "new_nms_gpu", "nms_func", "softnms_gpu", "softnms", "scale_by_iou",
]


# ---------------------------------------------------------------- helpers
def check_numpy_to_torch(x):
    """pcdet/utils/common_utils.py:15-18."""
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).float(), True
    return x, False


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _device_for_host_call(what: str) -> torch.device:
    """The GPU that executes a host-signature (``_cpu``) entry point.

    The reference calls these from forked DataLoader workers (database_sampler.py:246-247, box_utils.py:86,
    augmentor_utils.py:149).  CUDA cannot be initialised in a child forked after the parent touched it, and this
    package has no host implementation to fall back to (by design) -- so fail loudly and say what to do."""
    bad_fork = getattr(torch.cuda, "_is_in_bad_fork", None)
    if bad_fork is not None and bad_fork():
        raise RuntimeError(
            f"glenet_b200.{what} executes on the GPU and was called in a process forked after CUDA was initialised "
            "(a DataLoader worker?).  Use multiprocessing_context='spawn' for the workers, or keep the reference's host "
            "function for this call site: glenet_b200.shim.install() does so by default (cpu_entry_points=False).")
    if not torch.cuda.is_available():
        raise RuntimeError(f"glenet_b200.{what} needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _check_cuda_f32(t: torch.Tensor, name: str) -> None:
    # the reference's CHECK_INPUT exits the process on a CPU tensor (iou3d_nms.cpp:14-26) and
    # .data<float>() throws on another dtype; raise instead
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")


def _pairwise(fn_name: str, boxes_a: torch.Tensor, boxes_b: torch.Tensor) -> torch.Tensor:
    _check_cuda_f32(boxes_a, "boxes_a")
    _check_cuda_f32(boxes_b, "boxes_b")
    if boxes_a.device != boxes_b.device:
        raise RuntimeError("boxes_a and boxes_b must be on the same device")
    a = boxes_a.contiguous()
    b = boxes_b.contiguous()
    na, nb = a.shape[0], b.shape[0]
    out = torch.empty((na, nb), dtype=torch.float32, device=a.device)
    if na and nb:
        lib = _lib.load()
        with torch.cuda.device(a.device):
            rc = getattr(lib, fn_name)(a.data_ptr(), na, b.data_ptr(), nb, out.data_ptr(), _stream(a.device))
        _lib.check(rc, fn_name)
    return out


# ---------------------------------------------------------------- IoU
def _bev_iou_cpu_dialect_on_device(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """(N, M) BEV IoU of float32 CPU boxes in the CPU dialect, computed and LEFT on the GPU."""
    na, nb = a.shape[0], b.shape[0]
    dev = _device_for_host_call("boxes_bev_iou_cpu")
    lib = _lib.load()
    # one host buffer: [boxes_a | boxes_b | trig_a | trig_b] -> one H2D copy
    # (the trig tables are read as float4 => their offsets are padded to 16 bytes)
    o_b = na * 7
    o_ta = (o_b + nb * 7 + 3) // 4 * 4
    o_tb = o_ta + na * 4
    host = torch.empty(o_tb + nb * 4, dtype=torch.float32).pin_memory()
    host[:o_b].copy_(a.reshape(-1))
    host[o_b:o_b + nb * 7].copy_(b.reshape(-1))
    base = host.data_ptr()
    lib.glenet_host_trig4(a.data_ptr(), na, base + 4 * o_ta)
    lib.glenet_host_trig4(b.data_ptr(), nb, base + 4 * o_tb)
    d = host.to(dev, non_blocking=True)
    out = torch.empty((na, nb), dtype=torch.float32, device=dev)
    p = d.data_ptr()
    with torch.cuda.device(dev):
        rc = lib.glenet_boxes_iou_bev_cpu_dialect(p, p + 4 * o_ta, na, p + 4 * o_b, p + 4 * o_tb, nb,
                                                  out.data_ptr(), _stream(dev))
    _lib.check(rc, "glenet_boxes_iou_bev_cpu_dialect")
    out.record_stream(torch.cuda.current_stream(dev))
    return out


def boxes_bev_iou_cpu(boxes_a, boxes_b):
    """
    Args:
        boxes_a: (N, 7) [x, y, z, dx, dy, dz, heading]
        boxes_b: (M, 7) [x, y, z, dx, dy, dz, heading]

    Returns:
        ans_iou: (N, M)

    Reference: iou3d_nms_utils.py:52-68 -> boxes_iou_bev_cpu (iou3d_cpu.cpp:232-252), a
    single-threaded host loop.  Same signature (CPU tensors or numpy in, same kind out; the
    numpy flag follows ``boxes_b`` as in the reference), but the N x M clipping runs on the
    GPU in the CPU dialect: host-libm trig per box, no FMA contraction.
    """
    boxes_a, is_numpy = check_numpy_to_torch(boxes_a)
    boxes_b, is_numpy = check_numpy_to_torch(boxes_b)
    assert not (boxes_a.is_cuda or boxes_b.is_cuda), 'Only support CPU tensors'
    assert boxes_a.shape[1] == 7 and boxes_b.shape[1] == 7
    if boxes_a.dtype != torch.float32 or boxes_b.dtype != torch.float32:
        raise RuntimeError("boxes must be float32")   # reference: .data<float>() throws
    a = boxes_a.contiguous()
    b = boxes_b.contiguous()
    na, nb = a.shape[0], b.shape[0]
    ans_iou = boxes_a.new_zeros(torch.Size((na, nb)))
    if na and nb:
        ans_iou.copy_(_bev_iou_cpu_dialect_on_device(a, b))   # D2H, synchronising
    return ans_iou.numpy() if is_numpy else ans_iou


def boxes_iou_bev(boxes_a, boxes_b):
    """
    Args:
        boxes_a: (N, 7) [x, y, z, dx, dy, dz, heading]
        boxes_b: (M, 7) [x, y, z, dx, dy, dz, heading]

    Returns:
        ans_iou: (N, M)

    Reference: iou3d_nms_utils.py:71-85 -> boxes_iou_bev_gpu (iou3d_nms.cpp:70-88).
    """
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    return _pairwise("glenet_boxes_iou_bev_gpu", boxes_a, boxes_b)


def boxes_overlap_bev(boxes_a, boxes_b):
    """BEV overlap area (N, M): the native call inside the reference's boxes_iou3d_gpu
    (iou3d_nms_utils.py:106-107 -> boxes_overlap_bev_gpu, iou3d_nms.cpp:49-68)."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    return _pairwise("glenet_boxes_overlap_bev_gpu", boxes_a, boxes_b)


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """
    Args:
        boxes_a: (N, 7) [x, y, z, dx, dy, dz, heading]
        boxes_b: (M, 7) [x, y, z, dx, dy, dz, heading]

    Returns:
        ans_iou: (N, M)

    Reference: iou3d_nms_utils.py:88-121 (one custom kernel + ~10 torch elementwise kernels,
    each separately rounded).  Here: one fused kernel with the same rounding steps.
    """
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    return _pairwise("glenet_boxes_iou3d_gpu", boxes_a, boxes_b)


def _frames(mode: int, boxes_a: torch.Tensor, boxes_b: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    _check_cuda_f32(boxes_a, "boxes_a")
    _check_cuda_f32(boxes_b, "boxes_b")
    if boxes_a.device != boxes_b.device:
        raise RuntimeError("boxes_a and boxes_b must be on the same device")
    assert boxes_b.dim() == 3 and boxes_b.shape[2] == 7, "boxes_b must be (F, M, 7)"
    assert boxes_a.shape[-1] == 7 and boxes_a.dim() in (2, 3), "boxes_a must be (N, 7) or (F, N, 7)"
    frames, nb = boxes_b.shape[0], boxes_b.shape[1]
    a, b = boxes_a.contiguous(), boxes_b.contiguous()
    if a.dim() == 3:
        assert a.shape[0] == frames, "boxes_a and boxes_b disagree on the number of frames"
        na, stride_a = a.shape[1], a.shape[1] * 7
    else:
        na, stride_a = a.shape[0], 0
    if out is None:
        out = torch.empty((frames, na, nb), dtype=torch.float32, device=a.device)
    else:
        assert out.shape == (frames, na, nb) and out.dtype == torch.float32 and out.is_contiguous() and out.device == a.device
    if frames and na and nb:
        lib = _lib.load()
        with torch.cuda.device(a.device):
            rc = lib.glenet_boxes_iou_frames_gpu(mode, a.data_ptr(), stride_a, na, b.data_ptr(), nb * 7, nb, out.data_ptr(), frames, _stream(a.device))
        _lib.check(rc, "glenet_boxes_iou_frames_gpu")
    return out


def boxes_iou_bev_frames(boxes_a, boxes_b, out=None):
    """BEV IoU of every frame of a batch in ONE launch.

    Args:
        boxes_a: (N, 7) shared by all frames (e.g. the anchors) or (F, N, 7)
        boxes_b: (F, M, 7) e.g. the zero-padded ``gt_boxes`` of a batch
    Returns:
        (F, N, M); frame f equals ``boxes_iou_bev(boxes_a[f], boxes_b[f])`` bit for bit.

    Additive API.  The reference loops over the batch in Python and launches one kernel per frame
    (axis_aligned_target_assigner.py:60-105, proposal_target_layer.py:116-160); with one grid for all
    frames the tiles of different frames overlap on the SMs."""
    return _frames(1, boxes_a, boxes_b, out)


def boxes_overlap_bev_frames(boxes_a, boxes_b, out=None):
    """Frame-batched :func:`boxes_overlap_bev`."""
    return _frames(0, boxes_a, boxes_b, out)


def boxes_iou3d_gpu_frames(boxes_a, boxes_b, out=None):
    """Frame-batched :func:`boxes_iou3d_gpu`: (N, 7) or (F, N, 7) x (F, M, 7) -> (F, N, M)."""
    return _frames(2, boxes_a, boxes_b, out)


_MODES = {"overlap": 0, "bev": 1, "3d": 2}


def boxes_iou_frames_sparse(boxes_a, boxes_b, mode="bev", cap=None):
    """The non-zero elements of the frame-batched IoU matrix as a coordinate list -- the (F, N, M) matrix itself is
    never written.

    Args:
        boxes_a: (N, 7) shared by all frames or (F, N, 7);  boxes_b: (F, M, 7) or (M, 7) for a single frame
        mode: "bev" (boxes_iou_bev), "3d" (boxes_iou3d_gpu) or "overlap" (boxes_overlap_bev)
        cap: initial capacity of the list (grown and re-run automatically if too small)
    Returns:
        idx (K,) int64 flat indices ``f * N * M + row * M + col`` and val (K,) float32, in no particular order;
        ``val[k]`` is bit-identical to the dense call's element, every other element of the dense matrix is +0.0.

    Additive API for the reference's consumers that only reduce the matrix (see :func:`iou_max_overlaps_frames`)."""
    _check_cuda_f32(boxes_a, "boxes_a")
    _check_cuda_f32(boxes_b, "boxes_b")
    if boxes_b.dim() == 2:
        boxes_b = boxes_b.unsqueeze(0)
    assert boxes_b.dim() == 3 and boxes_b.shape[2] == 7 and boxes_a.shape[-1] == 7 and boxes_a.dim() in (2, 3)
    frames, nb = boxes_b.shape[0], boxes_b.shape[1]
    a, b = boxes_a.contiguous(), boxes_b.contiguous()
    if a.dim() == 3:
        assert a.shape[0] == frames
        na, stride_a = a.shape[1], a.shape[1] * 7
    else:
        na, stride_a = a.shape[0], 0
    dev = a.device
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    if cap is None:
        cap = max(1 << 16, 32 * frames * (na + nb))
    cap = int(min(cap, max(1, frames * na * nb)))
    lib = _lib.load()
    while True:
        idx = torch.empty((cap,), dtype=torch.int64, device=dev)
        val = torch.empty((cap,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.glenet_boxes_iou_frames_sparse_gpu(_MODES[mode], a.data_ptr(), stride_a, na, b.data_ptr(), nb * 7, nb, frames,
                                                        idx.data_ptr(), val.data_ptr(), cap, count.data_ptr(), _stream(dev))
        _lib.check(rc, "glenet_boxes_iou_frames_sparse_gpu")
        k = int(count.item())      # the list's length is data dependent: one sync, like the reference's nonzero() calls
        if k <= cap:
            return idx[:k], val[:k]
        cap = k


def iou_max_overlaps_frames(boxes_a, boxes_b, mode="bev"):
    """Row / column maxima of the IoU matrix of every frame without materialising it.

    Returns ``(a_max, a_argmax, b_max, b_argmax)`` with shapes (F, N), (F, N), (F, M), (F, M): exactly
    ``iou.max(dim=2)``, ``iou.argmax(dim=2)`` (first index among ties, numpy's rule -- the reference goes through
    numpy, axis_aligned_target_assigner.py:141-150), ``iou.max(dim=1)``, ``iou.argmax(dim=1)`` of
    ``iou = boxes_iou_*_frames(boxes_a, boxes_b)``.  Rows / columns without any overlap report max 0, argmax 0.
    These four vectors are what the anchor assignment and the RoI sampler consume
    (axis_aligned_target_assigner.py:141-165, proposal_target_layer.py:113-114).  Two launches -- the IoU kernel (atomic max
    on packed (value, index) keys) and one decode kernel -- and no host sync."""
    _check_cuda_f32(boxes_a, "boxes_a")
    _check_cuda_f32(boxes_b, "boxes_b")
    if boxes_b.dim() == 2:
        boxes_b = boxes_b.unsqueeze(0)
    assert boxes_b.dim() == 3 and boxes_b.shape[2] == 7 and boxes_a.shape[-1] == 7 and boxes_a.dim() in (2, 3)
    frames, nb = boxes_b.shape[0], boxes_b.shape[1]
    a, b = boxes_a.contiguous(), boxes_b.contiguous()
    if a.dim() == 3:
        assert a.shape[0] == frames
        na, stride_a = a.shape[1], a.shape[1] * 7
    else:
        na, stride_a = a.shape[0], 0
    dev = a.device
    row_key = torch.empty((frames, na), dtype=torch.int64, device=dev)
    col_key = torch.empty((frames, nb), dtype=torch.int64, device=dev)
    a_max = torch.empty((frames, na), dtype=torch.float32, device=dev)
    a_arg = torch.empty((frames, na), dtype=torch.int64, device=dev)
    b_max = torch.empty((frames, nb), dtype=torch.float32, device=dev)
    b_arg = torch.empty((frames, nb), dtype=torch.int64, device=dev)
    if frames:
        lib = _lib.load()
        with torch.cuda.device(dev):
            rc = lib.glenet_boxes_iou_frames_max_gpu(_MODES[mode], a.data_ptr(), stride_a, na, b.data_ptr(), nb * 7, nb, frames,
                                                     row_key.data_ptr(), col_key.data_ptr(), _stream(dev))
            _lib.check(rc, "glenet_boxes_iou_frames_max_gpu")
            # key = (float bits << 32) | (0xffffffff - index); key == 0 <=> nothing non-zero on that row / column
            rc = lib.glenet_iou_keys_decode_gpu(row_key.data_ptr(), frames * na, col_key.data_ptr(), frames * nb,
                                                a_max.data_ptr(), a_arg.data_ptr(), b_max.data_ptr(), b_arg.data_ptr(), _stream(dev))
        _lib.check(rc, "glenet_iou_keys_decode_gpu")
    return a_max, a_arg, b_max, b_arg


def _aligned(mode: int, boxes_a: torch.Tensor, boxes_b: torch.Tensor, group: int) -> torch.Tensor:
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    _check_cuda_f32(boxes_a, "boxes_a")
    _check_cuda_f32(boxes_b, "boxes_b")
    a, b = boxes_a.contiguous(), boxes_b.contiguous()
    na = a.shape[0]
    assert group >= 1 and b.shape[0] * group >= na, "boxes_b must hold ceil(N / group) rows"
    out = torch.empty((na,), dtype=torch.float32, device=a.device)
    if na:
        lib = _lib.load()
        with torch.cuda.device(a.device):
            rc = lib.glenet_boxes_iou_aligned_gpu(mode, a.data_ptr(), na, b.data_ptr(), group, out.data_ptr(), _stream(a.device))
        _lib.check(rc, "glenet_boxes_iou_aligned_gpu")
    return out


def boxes_iou_bev_aligned(boxes_a, boxes_b, group=1):
    """out[i] = BEV IoU(boxes_a[i], boxes_b[i // group]).  Additive API: the block-diagonal of
    boxes_iou_bev for the CVAE label-uncertainty workload (``group`` sampled boxes per GT)."""
    return _aligned(1, boxes_a, boxes_b, group)


def boxes_iou3d_aligned(boxes_a, boxes_b, group=1):
    """out[i] = 3D IoU(boxes_a[i], boxes_b[i // group]); same arithmetic as boxes_iou3d_gpu."""
    return _aligned(2, boxes_a, boxes_b, group)


# ---------------------------------------------------------------- NMS
def _nms_sorted(fn_name: str, boxes_sorted: torch.Tensor, thresh: float):
    """boxes_sorted: (F, n, 7) contiguous CUDA float32.  Returns (keep (F, n) int64, num (F,) int32), on device."""
    frames, n = boxes_sorted.shape[0], boxes_sorted.shape[1]
    dev = boxes_sorted.device
    keep = torch.empty((frames, n), dtype=torch.int64, device=dev)
    num = torch.zeros((frames,), dtype=torch.int32, device=dev)
    if frames and n:
        lib = _lib.load()
        ws_bytes = lib.glenet_nms_workspace_bytes(frames, n)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = getattr(lib, fn_name)(boxes_sorted.data_ptr(), frames, n, float(thresh), keep.data_ptr(),
                                       num.data_ptr(), ws.data_ptr(), ws_bytes, _stream(dev))
        _lib.check(rc, fn_name)
    return keep, num


def _nms(fn_name: str, boxes, scores, thresh, pre_maxsize=None):
    assert boxes.shape[1] == 7
    _check_cuda_f32(boxes, "boxes")
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    keep, num = _nms_sorted(fn_name, boxes.unsqueeze(0), thresh)
    num_out = int(num.item())   # the one unavoidable sync: the result length
    return order[keep[0, :num_out]].contiguous(), None


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """
    :param boxes: (N, 7) [x, y, z, dx, dy, dz, heading]
    :param scores: (N)
    :param thresh:
    :return: (indices into ``boxes`` of the kept boxes, by descending score; None)

    Reference: iou3d_nms_utils.py:182-197 -> nms_gpu (iou3d_nms.cpp:90-136).  ``**kwargs``
    swallows the NMS_CONFIG dict that model_nms_utils.py:50-52 splats into the call.
    """
    return _nms("glenet_nms_gpu", boxes, scores, thresh, pre_maxsize)


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    """
    :param boxes: (N, 7) [x, y, z, dx, dy, dz, heading]
    :param scores: (N)
    :param thresh:
    :return:

    Reference: iou3d_nms_utils.py:276-290 -> nms_normal_gpu (iou3d_nms.cpp:139-186).
    """
    return _nms("glenet_nms_normal_gpu", boxes, scores, thresh)


def _nms_batch(fn_name: str, boxes, scores, thresh):
    assert boxes.dim() == 3 and boxes.shape[2] == 7 and scores.shape == boxes.shape[:2]
    _check_cuda_f32(boxes, "boxes")
    order = scores.sort(1, descending=True)[1]
    sorted_boxes = torch.gather(boxes, 1, order.unsqueeze(-1).expand(-1, -1, 7)).contiguous()
    keep, num = _nms_sorted(fn_name, sorted_boxes, thresh)
    # map sorted positions back to indices of the caller's tensors; rows are valid up to num[f]
    return torch.gather(order, 1, keep.clamp_(0, max(boxes.shape[1] - 1, 0))), num


def nms_gpu_batch(boxes, scores, thresh):
    """Additive API: NMS of F independent frames in one launch pair, no host synchronisation.
    boxes (F, N, 7), scores (F, N) -> (keep (F, N) int64, num_keep (F,) int32); row f is valid
    up to num_keep[f] and equals nms_gpu(boxes[f], scores[f], thresh)[0]."""
    return _nms_batch("glenet_nms_gpu", boxes, scores, thresh)


def nms_normal_gpu_batch(boxes, scores, thresh):
    """Batched nms_normal_gpu, see nms_gpu_batch."""
    return _nms_batch("glenet_nms_normal_gpu", boxes, scores, thresh)


# GLENet's variance-voting NMS / soft-NMS (iou3d_nms_utils.py:200-356): host control flow over the IoU functions above
from .variance_nms import new_nms_gpu, nms_func, scale_by_iou, softnms, softnms_gpu  # noqa: E402,F401
