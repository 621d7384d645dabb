"""Multi-GPU exchange (glenet_b200/sharded.py, csrc/exchange.cuh) on real devices: ``pytest -m gpu``.

Two ranks, one process each.  With >= 2 GPUs the ranks sit on cuda:0 / cuda:1 and rendezvous over NCCL (the same wiring as
bench.py under torchrun), and the ``torch.distributed`` formulation (``boxes_iou_sharded``) is checked over NCCL as well.
On a one-GPU box both ranks share cuda:0 and the IPC handles travel over gloo: the exchange protocol (peer mappings,
system-scope atomics, flags, double buffering) is the same code, only time-sliced instead of concurrent.
"""
import os
import socket

import numpy as np
import pytest
import torch

from glenet_b200 import iou3d_nms_utils as I
from glenet_b200 import sharded, synth

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    anchors = synth.anchors_kitti3()[:30000:1].contiguous()             # 30 000 rows: slabs of 15 040 / 14 960
    gts = torch.stack([synth.kitti_boxes(100, 300 + f) for f in range(3)])
    gts[1, 60:] = 0                                                        # zero padding rows, as the assigner passes them
    return anchors, gts


def test_single_rank_exchange_equals_dense_and_max_overlaps():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda:0")
    anchors, gts = _problem()
    a, g = anchors.to(dev), gts.to(dev)
    dense = I.boxes_iou_bev_frames(a, g)
    win = sharded.ExchangeWindow(frames=3, nb=100, list_cap=1 << 20)
    try:
        for _ in range(3):                                                 # key buffers are re-zeroed by the decode kernel: repeatable
            res = sharded.anchor_assign_sharded(a, g, win)
            assert res["rows"] == (0, 30000) and torch.equal(res["iou"], dense)
            assert torch.equal(res["row_max"], dense.max(dim=2).values) and torch.equal(res["col_max"], dense.max(dim=1).values)
            assert np.array_equal(res["row_argmax"].cpu().numpy(), dense.cpu().numpy().argmax(axis=2))
            assert np.array_equal(res["col_argmax"].cpu().numpy(), dense.cpu().numpy().argmax(axis=1))
            r2 = sharded.anchor_assign_sharded(a, g, win, dense=False)
            assert r2["iou"] is None and torch.equal(r2["col_max"], dense.max(dim=1).values)
            full = sharded.boxes_iou_gather_sharded(a, g, win)
            assert torch.equal(full, dense)
            full2 = sharded.boxes_iou_gather_sharded(a, g, win, fill_stream=torch.cuda.Stream(dev))
            assert torch.equal(full2, dense)
        # matrix-free maxima through the same decode kernel
        a_max, a_arg, b_max, b_arg = I.iou_max_overlaps_frames(a, g)
        assert torch.equal(a_max, dense.max(dim=2).values) and np.array_equal(b_arg.cpu().numpy(), dense.cpu().numpy().argmax(axis=1))
        assert win.status() == 0
        # a list that is too small is reported, not silently truncated
        small = sharded.ExchangeWindow(frames=3, nb=100, list_cap=64)
        sharded.boxes_iou_gather_sharded(a, g, small)
        assert small.status() & 2
        small.close()
    finally:
        win.close()


def _worker(rank, world, port, shared_gpu, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dev = torch.device("cuda", 0 if shared_gpu else rank)
        torch.cuda.set_device(dev)
        if shared_gpu:
            dist.init_process_group("gloo", rank=rank, world_size=world)
        else:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        anchors, gts = _problem()
        a, g = anchors.to(dev), gts.to(dev)
        dense = I.boxes_iou_bev_frames(a, g)                               # the single-GPU answer, computed by every rank for itself
        want_cmax, want_carg = dense.max(dim=1).values, torch.from_numpy(dense.cpu().numpy().argmax(axis=1)).to(dev)
        start, stop = sharded.shard_rows(a.shape[0], world, rank)
        win = sharded.ExchangeWindow(frames=3, nb=100, list_cap=1 << 19)
        out = {}
        # several steps back to back without any host synchronisation in between (exercises the double buffering), with a
        # different problem every other step so that stale keys / list entries would show
        g_alt = g.flip(0).contiguous()
        dense_alt = dense.flip(0)
        results = []
        for step in range(6):
            gg = g if step % 2 == 0 else g_alt
            res = sharded.anchor_assign_sharded(a, gg, win)
            results.append({k: (v.clone() if torch.is_tensor(v) else v) for k, v in res.items()})
        fulls = [sharded.boxes_iou_gather_sharded(a, g if step % 2 == 0 else g_alt, win).clone() for step in range(4)]
        torch.cuda.synchronize()
        ok = True
        for step, res in enumerate(results):
            d = dense if step % 2 == 0 else dense_alt
            ok &= res["rows"] == (start, stop)
            ok &= torch.equal(res["iou"], d[:, start:stop])
            ok &= torch.equal(res["row_max"], d[:, start:stop].max(dim=2).values)
            ok &= bool(np.array_equal(res["row_argmax"].cpu().numpy(), d[:, start:stop].cpu().numpy().argmax(axis=2)))
            ok &= torch.equal(res["col_max"], d.max(dim=1).values)
            ok &= bool(np.array_equal(res["col_argmax"].cpu().numpy(), d.cpu().numpy().argmax(axis=1)))
        out["assign"] = bool(ok)
        out["gather"] = all(torch.equal(f, dense if step % 2 == 0 else dense_alt) for step, f in enumerate(fulls))
        out["status"] = win.status()
        if not shared_gpu:                                                  # the torch.distributed formulation over NCCL
            full = sharded.boxes_iou_sharded(a, g[0], gather="full")
            red = sharded.boxes_iou_sharded(a, g[0], gather="reductions")
            out["nccl_full"] = torch.equal(full, dense[0])
            out["nccl_red"] = torch.equal(red["col_max"], want_cmax[0]) and torch.equal(red["col_argmax"], want_carg[0]) and \
                torch.equal(red["row_max"], dense[0, start:stop].max(dim=1).values)
        win.close()
        dist.destroy_process_group()
        q.put((rank, out))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, {"error": f"{e!r}\n{traceback.format_exc()}"}))


def test_two_rank_exchange_in_kernel_and_nccl():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import torch.multiprocessing as mp
    shared_gpu = torch.cuda.device_count() < 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shared_gpu, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in procs:
        rank, out = q.get(timeout=600)
        got[rank] = out
    for p in procs:
        p.join(timeout=60)
    for rank in (0, 1):
        assert "error" not in got[rank], got[rank].get("error")
        assert got[rank]["assign"] and got[rank]["gather"] and got[rank]["status"] == 0, (rank, got[rank])
        if not shared_gpu:
            assert got[rank]["nccl_full"] and got[rank]["nccl_red"], (rank, got[rank])
