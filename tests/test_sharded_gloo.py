"""world_size-2 gloo tests (CPU) of the row-sharding plumbing; the compute is the C oracle here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from glenet_b200 import sharded, synth


def test_shard_rows_cover_and_align():
    for n in (0, 1, 63, 64, 65, 1000, 211200):
        for world in (1, 2, 4, 8):
            spans = [sharded.shard_rows(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
                assert e0 == s1 and s0 <= e0
            assert all(s % 64 == 0 for s, _ in spans if s < n)
    with pytest.raises(ValueError):
        sharded.shard_rows(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import capi

    def compute(a, b):
        return torch.from_numpy(capi.boxes_iou_bev(a, b, dialect=capi.GPU))

    a, _ = synth.proposals(333, 9, 0)
    b = synth.kitti_boxes(17, 0)
    b[:9] = synth.kitti_boxes(9, 0)      # the proposals' centres => plenty of overlaps
    full_ref = compute(a, b)
    slab, (s, e) = sharded.boxes_iou_sharded(a, b, compute=compute)
    ok = torch.equal(slab, full_ref[s:e])
    full = sharded.boxes_iou_sharded(a, b, gather="full", compute=compute)
    ok &= torch.equal(full, full_ref)
    red = sharded.boxes_iou_sharded(a, b, gather="reductions", compute=compute)
    ok &= torch.equal(red["row_max"], full_ref[s:e].max(1)[0])
    ok &= torch.equal(red["col_max"], full_ref.max(0)[0])
    # smallest row index among equal maxima
    want_arg = torch.tensor([int((full_ref[:, j] == full_ref[:, j].max()).nonzero()[0]) for j in range(b.shape[0])])
    ok &= torch.equal(red["col_argmax"], want_arg)
    q.put((rank, bool(ok), (s, e)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_iou():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == (0, 192) and res[1][2] == (192, 333)
