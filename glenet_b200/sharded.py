"""Row-sharded pairwise IoU over the GPUs of one NVLink / NVSwitch box (one process per GPU).

Every element of the (N, M) IoU matrix depends on one (row, column) pair only, so ``boxes_a`` is split into contiguous
64-row-aligned slabs, ``boxes_b`` (<= a few hundred boxes) is replicated and each rank computes its slab.  What a consumer
needs from the OTHER ranks is small, and here it moves inside the IoU kernel itself -- peer stores and system-scope atomics
into CUDA-IPC-mapped "exchange windows" (``csrc/exchange.cuh``) -- not through NCCL:

* :func:`anchor_assign_sharded` -- the slab stays on its GPU, plus the reductions the anchor assigner consumes
  (``axis_aligned_target_assigner.py:141-165``): row max / argmax of the local rows and column max / first row over ALL rows;
* :func:`boxes_iou_gather_sharded` -- the whole matrix on every rank: each rank zero-fills its own copy and only the non-zero
  elements (< 1 % of an anchor sweep) cross NVLink as a coordinate list.

:func:`boxes_iou_sharded` is the plain ``torch.distributed`` formulation of the same results (``all_reduce`` of the
reductions, ``all_gather_into_tensor`` of the dense slabs -- NVLink-bound for the full matrix); it is what the fused paths
are measured against, and, with an injectable compute function, what the CPU / gloo tests exercise.

The reference never shards geometry ops; this is additive API.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

ROW_ALIGN = 64
_MODES = {"overlap": 0, "bev": 1, "3d": 2}


def shard_rows(n: int, world: int, rank: int, align: int = ROW_ALIGN) -> Tuple[int, int]:
    """Contiguous slab [start, stop) of rank ``rank``; slab size is a multiple of ``align`` (last may be short/empty)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = slab_rows(n, world, align)
    start = min(n, rank * per)
    return start, min(n, start + per)


def slab_rows(n: int, world: int, align: int = ROW_ALIGN) -> int:
    per = (n + world - 1) // world
    return (per + align - 1) // align * align


def _world_rank(group) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


# ------------------------------------------------------------------ exchange windows (CUDA IPC)
class ExchangeWindow:
    """This rank's exchange window plus the IPC mappings of every peer's window.

    Collective: every rank of ``group`` constructs it with the same ``(frames, nb, list_cap)`` -- the largest problem it
    will carry: ``frames * nb`` column keys and coordinate lists of ``list_cap`` entries per source rank (0 = no gather).
    The 64-byte IPC handles travel through ``torch.distributed``; afterwards no collective of that library is on the data
    path.  ``close()`` unmaps and frees (also collective in effect: peers must not use the window afterwards)."""

    def __init__(self, frames: int, nb: int, list_cap: int = 0, group=None, device: Optional[torch.device] = None):
        from . import _lib
        self._lib = lib = _lib.load()
        self.frames, self.nb, self.list_cap, self.group = int(frames), int(nb), int(list_cap), group
        self.world, self.rank = _world_rank(group)
        if self.world > 8:
            raise ValueError("an exchange spans the GPUs of one box (<= 8 ranks)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.bytes = int(lib.glenet_exchange_window_bytes(self.frames, self.nb, self.list_cap))
        self.assign_step = 0
        self.gather_step = 0
        self._scratch = {}
        self._peers = []
        with torch.cuda.device(self.device):
            local = ctypes.c_void_p()
            _lib.check(lib.glenet_symm_alloc(self.bytes, ctypes.byref(local)), "glenet_symm_alloc")
            self.local = local.value
            ptrs = [None] * self.world
            ptrs[self.rank] = self.local
            if self.world > 1:
                handle = ctypes.create_string_buffer(64)
                _lib.check(lib.glenet_symm_export(self.local, handle), "glenet_symm_export")
                mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).clone()
                backend = dist.get_backend(group)
                mine = mine.to(self.device) if backend == "nccl" else mine
                everyone = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(everyone, mine, group=group)
                for p, h in enumerate(everyone):
                    if p == self.rank:
                        continue
                    peer = ctypes.c_void_p()
                    raw = bytes(h.cpu().numpy().tobytes())
                    _lib.check(lib.glenet_symm_import(raw, ctypes.byref(peer)), f"glenet_symm_import(rank {p})")
                    ptrs[p] = peer.value
                    self._peers.append(peer.value)
                torch.cuda.synchronize(self.device)
                dist.barrier(group=group)      # every window exists, is zeroed and is mapped everywhere
        self.table = (ctypes.c_void_p * self.world)(*ptrs)

    def status(self) -> int:
        """Error bits left by the consumer kernels (synchronises): 1 = timed out waiting for a peer, 2 = list overflow."""
        out = ctypes.c_uint(0)
        from . import _lib
        with torch.cuda.device(self.device):
            _lib.check(self._lib.glenet_exchange_status(self.local, ctypes.byref(out)), "glenet_exchange_status")
        return int(out.value)

    def scratch(self, name: str, shape, dtype, zero: bool = False) -> torch.Tensor:
        key = (name, tuple(shape), dtype)
        t = self._scratch.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)
            self._scratch[key] = t
        return t

    def close(self) -> None:
        if getattr(self, "local", None) is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            if self.world > 1:
                dist.barrier(group=self.group)   # nobody is still writing into a window that is about to go away
            for p in self._peers:
                self._lib.glenet_symm_unmap(p)
            self._lib.glenet_symm_free(self.local)
        self._peers, self.local = [], None

    def __del__(self):   # best effort; explicit close() is the collective-safe way
        try:
            if getattr(self, "local", None) is not None and self.world == 1:
                self.close()
        except Exception:
            pass


def _prep(boxes_a: torch.Tensor, boxes_b: torch.Tensor, window: ExchangeWindow):
    if not (boxes_a.is_cuda and boxes_b.is_cuda) or boxes_a.dtype != torch.float32 or boxes_b.dtype != torch.float32:
        raise RuntimeError("boxes must be float32 CUDA tensors")
    if boxes_b.dim() == 2:
        boxes_b = boxes_b.unsqueeze(0)
    assert boxes_b.dim() == 3 and boxes_b.shape[2] == 7 and boxes_a.dim() == 2 and boxes_a.shape[1] == 7, \
        "boxes_a (N, 7) shared by all frames, boxes_b (F, M, 7) or (M, 7)"
    frames, nb = boxes_b.shape[0], boxes_b.shape[1]
    if frames * nb > window.frames * window.nb:
        raise ValueError("the exchange window was created for fewer column keys")
    n = boxes_a.shape[0]
    start, stop = shard_rows(n, window.world, window.rank)
    return boxes_a[start:stop].contiguous(), boxes_b.contiguous(), frames, nb, n, start, stop


def anchor_assign_sharded(boxes_a, boxes_b, window: ExchangeWindow, mode: str = "bev", dense: bool = True, out=None):
    """This rank's slab of IoU(boxes_a, boxes_b[f]) for every frame f, plus the assigner's reductions, exchanged in-kernel.

    Args:
        boxes_a: (N, 7), the same on every rank (e.g. the anchors); rank r works on rows ``shard_rows(N, world, r)``
        boxes_b: (F, M, 7) or (M, 7), the same on every rank (e.g. the padded GT boxes of a batch)
        dense: also write the slab of the matrix (False: reductions only, nothing is materialised)
    Returns a dict: ``rows`` (start, stop); ``iou`` (F, rows, M) or None; ``row_max`` / ``row_argmax`` (F, rows) -- the
    ``iou.max(dim=2)`` / first argmax of the local rows; ``col_max`` / ``col_argmax`` (F, M) -- max over the rows of ALL
    ranks and the smallest GLOBAL row attaining it (identical on every rank; rows / columns without overlap: 0, 0).
    No host synchronisation and no NCCL call; the scratch / output vectors belong to ``window`` and are reused by the next call
    unless ``out`` is given.  Consumer: ``axis_aligned_target_assigner.py:141-165``."""
    from . import _lib
    a, b, frames, nb, n, start, stop = _prep(boxes_a, boxes_b, window)
    rows = stop - start
    dev = a.device
    iou = None
    if dense:
        iou = out if out is not None else torch.empty((frames, rows, nb), dtype=torch.float32, device=dev)
        assert iou.shape == (frames, rows, nb) and iou.is_contiguous() and iou.dtype == torch.float32
    row_key = window.scratch("row_key", (frames, max(rows, 1)), torch.int64, zero=True)
    row_max = window.scratch("row_max", (frames, rows), torch.float32)
    row_arg = window.scratch("row_arg", (frames, rows), torch.int64)
    col_max = window.scratch("col_max", (frames, nb), torch.float32)
    col_arg = window.scratch("col_arg", (frames, nb), torch.int64)
    window.assign_step += 1
    with torch.cuda.device(dev):
        rc = window._lib.glenet_boxes_iou_frames_assign_gpu(
            _MODES[mode], a.data_ptr(), 0, rows, b.data_ptr(), nb * 7, nb, frames, iou.data_ptr() if dense else None, start, n,
            row_key.data_ptr(), row_max.data_ptr(), row_arg.data_ptr(), col_max.data_ptr(), col_arg.data_ptr(),
            window.world, window.rank, window.table, window.list_cap, window.assign_step & 0xffffffff,
            torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "glenet_boxes_iou_frames_assign_gpu")
    return {"rows": (start, stop), "iou": iou, "row_max": row_max, "row_argmax": row_arg, "col_max": col_max, "col_argmax": col_arg}


def boxes_iou_gather_sharded(boxes_a, boxes_b, window: ExchangeWindow, mode: str = "bev", out=None, fill_stream=None):
    """The full (F, N, M) IoU matrix on EVERY rank, each rank computing only its row slab.

    The zeros never travel: every rank zero-fills its own copy (on ``fill_stream`` if given, so that the fill overlaps the
    IoU kernel) and the non-zero elements of all slabs arrive as coordinate lists written by the peers' kernels.
    ``window.list_cap`` bounds the non-zero elements per rank and call (``window.status() & 2`` reports an overflow)."""
    from . import _lib
    a, b, frames, nb, n, start, stop = _prep(boxes_a, boxes_b, window)
    if window.list_cap <= 0:
        raise ValueError("the exchange window was created without coordinate lists (list_cap = 0)")
    dev = a.device
    full = out if out is not None else torch.empty((frames, n, nb), dtype=torch.float32, device=dev)
    assert full.shape == (frames, n, nb) and full.is_contiguous() and full.dtype == torch.float32
    main = torch.cuda.current_stream(dev)
    window.gather_step += 1

    def call(zero_fill):
        with torch.cuda.device(dev):
            rc = window._lib.glenet_boxes_iou_frames_gather_gpu(
                _MODES[mode], a.data_ptr(), 0, stop - start, b.data_ptr(), nb * 7, nb, frames, full.data_ptr(), zero_fill, start, n,
                window.world, window.rank, window.table, window.list_cap, window.gather_step & 0xffffffff, main.cuda_stream)
        _lib.check(rc, "glenet_boxes_iou_frames_gather_gpu")

    if fill_stream is None:
        call(1)
    else:
        fill_stream.wait_stream(main)              # the previous consumer of `full` is done
        with torch.cuda.stream(fill_stream):
            full.zero_()
        call(2)                                    # the IoU kernel does not touch `full`: it runs under the fill
        main.wait_stream(fill_stream)
        call(3)                                    # scatter: after the fill and after every rank's flag
    return full


# ------------------------------------------------------------------ the torch.distributed formulation (baseline, CPU-testable)
def _default_compute(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    from .iou3d_nms_utils import boxes_iou_bev
    return boxes_iou_bev(a, b)


def boxes_iou_sharded(boxes_a: torch.Tensor, boxes_b: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                      gather: Optional[str] = None, compute: Callable = _default_compute):
    """Compute this rank's slab of IoU(boxes_a, boxes_b) and exchange results with ``torch.distributed`` collectives.

    Returns ``(slab, (start, stop))`` for ``gather=None``; the full (N, M) matrix for ``gather="full"``
    (``all_gather_into_tensor`` of the dense slabs: 1.35 GB for the anchor sweep, NVLink-bound); a dict of assigner
    reductions for ``gather="reductions"``: ``row_max``/``row_argmax`` for the local rows and ``col_max``/``col_argmax``
    over ALL rows (smallest row index among equal maxima; ``all_reduce`` MAX then MIN)."""
    world, rank = _world_rank(group)
    n, m = boxes_a.shape[0], boxes_b.shape[0]
    start, stop = shard_rows(n, world, rank)
    slab = compute(boxes_a[start:stop].contiguous(), boxes_b)
    if gather is None:
        return slab, (start, stop)
    if gather == "full":
        if world == 1:
            return slab
        per = slab_rows(n, world)
        padded = slab.new_zeros((per, m))
        padded[: stop - start] = slab
        full = slab.new_empty((per * world, m))
        dist.all_gather_into_tensor(full, padded, group=group)
        return full[:n]
    if gather == "reductions":
        if stop > start and m > 0:
            row_max, row_argmax = slab.max(dim=1)
            col_max, col_arg_local = slab.max(dim=0)
            col_argmax = col_arg_local + start
        else:
            row_max = slab.new_zeros((stop - start,))
            row_argmax = torch.zeros((stop - start,), dtype=torch.int64, device=slab.device)
            col_max = slab.new_full((m,), -1.0)
            col_argmax = torch.full((m,), n, dtype=torch.int64, device=slab.device)
        if world > 1:
            gmax = col_max.clone()
            dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
            cand = torch.where(col_max == gmax, col_argmax, torch.full_like(col_argmax, n))
            dist.all_reduce(cand, op=dist.ReduceOp.MIN, group=group)
            col_max, col_argmax = gmax, cand
        return {"rows": (start, stop), "row_max": row_max, "row_argmax": row_argmax, "col_max": col_max, "col_argmax": col_argmax}
    raise ValueError(f"unknown gather mode {gather!r}")
