"""Row-sharded pairwise IoU over the GPUs of one NVLink / NVSwitch box (one process per GPU).

Every element of the (N, M) IoU matrix depends on one (row, column) pair only, so ``boxes_a`` is
split into contiguous 64-row-aligned slabs, ``boxes_b`` (<= a few hundred boxes) is replicated, and
each rank computes its slab with no data-path collective.  What moves over NCCL afterwards is only
what a consumer asks for:

* ``gather="reductions"``: the row/column max + argmax that the anchor assigner consumes
  (``axis_aligned_target_assigner.py:147-152``) -- O(N/world + M) values per rank;
* ``gather="full"``: the whole matrix on every rank (``all_gather_into_tensor`` of the slabs);
  1.35 GB for the anchor sweep, i.e. NVLink-bound -- provided for completeness, not the fast path.

The reference never shards geometry ops; this is additive API.
The compute callable is injectable so that the plumbing is testable on CPU with gloo.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

ROW_ALIGN = 64


def shard_rows(n: int, world: int, rank: int, align: int = ROW_ALIGN) -> Tuple[int, int]:
    """Contiguous slab [start, stop) of rank ``rank``; slab size is a multiple of ``align`` (last may be short/empty)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = (n + world - 1) // world
    per = (per + align - 1) // align * align
    start = min(n, rank * per)
    return start, min(n, start + per)


def slab_rows(n: int, world: int, align: int = ROW_ALIGN) -> int:
    per = (n + world - 1) // world
    return (per + align - 1) // align * align


def _default_compute(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    from .iou3d_nms_utils import boxes_iou_bev
    return boxes_iou_bev(a, b)


def boxes_iou_sharded(boxes_a: torch.Tensor, boxes_b: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                      gather: Optional[str] = None, compute: Callable = _default_compute):
    """Compute this rank's slab of IoU(boxes_a, boxes_b).

    Returns ``(slab, (start, stop))`` for ``gather=None``; the full (N, M) matrix for ``gather="full"``;
    a dict of assigner reductions for ``gather="reductions"``:
    ``row_max``/``row_argmax`` for the local rows and ``col_max``/``col_argmax`` over ALL rows
    (smallest row index among equal maxima).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
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
