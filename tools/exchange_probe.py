"""Single-GPU timings that bound the row-sharded sweep: the per-rank kernels on the slab sizes of world = 1, 2, 4, 8
(the exchange itself is a few KB / a few MB on top).  python tools/exchange_probe.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import iou3d_nms_utils as I, sharded, synth  # noqa: E402

dev = torch.device("cuda:0")
anchors = synth.anchors_kitti3().to(dev)
gts = torch.stack([synth.kitti_boxes(100, 101 + f) for f in range(16)]).to(dev)


def ev(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3   # us


res = {}
win = sharded.ExchangeWindow(frames=16, nb=100, list_cap=1 << 22)
side = torch.cuda.Stream(dev)
for world in (1, 2, 4, 8):
    rows = sharded.slab_rows(anchors.shape[0], world)
    a = anchors[:rows].contiguous()
    out = torch.empty((16, rows, 100), dtype=torch.float32, device=dev)
    r = {"rows": rows}
    r["dense_us"] = ev(lambda: I.boxes_iou_bev_frames(a, gts, out=out))
    r["assign_dense_us"] = ev(lambda: sharded.anchor_assign_sharded(a, gts, win, out=out))
    r["assign_keys_only_us"] = ev(lambda: sharded.anchor_assign_sharded(a, gts, win, dense=False))
    res[f"world{world}_slab"] = r
full = torch.empty((16, anchors.shape[0], 100), dtype=torch.float32, device=dev)
res["gather_world1_us"] = ev(lambda: sharded.boxes_iou_gather_sharded(anchors, gts, win, out=full), 10)
res["gather_world1_fillstream_us"] = ev(lambda: sharded.boxes_iou_gather_sharded(anchors, gts, win, out=full, fill_stream=side), 10)
res["zero_fill_full_us"] = ev(lambda: full.zero_(), 10)
res["status"] = win.status()
win.close()
print(json.dumps(res, indent=1))
