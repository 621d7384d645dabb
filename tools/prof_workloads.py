"""Tiny driver for ncu captures: runs one named workload a few times.

    ncu --set full ... python tools/prof_workloads.py pib|iou_sparse|iou_frames|iou_dense|nms [iters]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import iou3d_nms_utils as I, roiaware_pool3d_utils as R, synth  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3

if what == "pib":
    B = 128
    boxes = torch.stack([synth.waymo_boxes(200, 100 + f) for f in range(B)]).to(dev)
    base = synth.points(180000, boxes[0].cpu(), synth.WAYMO_RANGE, 0.05, seed=5).to(dev)
    pts = (base.unsqueeze(0).repeat(B, 1, 1) + torch.randn(B, 180000, 3, device=dev) * 0.01).contiguous()
    fn = lambda: R.points_in_boxes_gpu(pts, boxes)
elif what == "iou_sparse":
    a = synth.anchors_kitti3().to(dev)
    b = synth.kitti_boxes(100, 4).to(dev)
    fn = lambda: I.boxes_iou_bev(a, b)
elif what == "iou_frames":   # the bench's step: 16 frames of the anchor sweep in one launch, 1.35 GB of result slabs
    a = synth.anchors_kitti3().to(dev)
    b = torch.stack([synth.kitti_boxes(100, 100 + f + 1) for f in range(16)]).to(dev)
    out = torch.empty((16, a.shape[0], 100), device=dev)
    fn = lambda: I.boxes_iou_bev_frames(a, b, out=out)
elif what == "iou_dense":
    s, g = synth.cvae_samples(20000, 30, 0)
    s, g = s.to(dev), g.to(dev)
    fn = lambda: I.boxes_iou3d_aligned(s, g, 30)
elif what == "iou_dense_pair":
    p, _ = synth.proposals(4096, 20, 0)
    p = p.to(dev)
    fn = lambda: I.boxes_iou_bev(p, p)
elif what == "nms":
    fb, fs = [], []
    for f in range(8):
        b, s = synth.proposals(4096, 20, 20 + f)
        fb.append(b); fs.append(s)
    fb, fs = torch.stack(fb).to(dev), torch.stack(fs).to(dev)
    fn = lambda: I.nms_gpu_batch(fb, fs, 0.7)
elif what == "assign":      # the bench's multi-GPU step on one rank: dense slab + assigner reductions (OUT_BOTH) + decode
    from glenet_b200 import sharded
    a = synth.anchors_kitti3().to(dev)
    b = torch.stack([synth.kitti_boxes(100, 100 + f + 1) for f in range(16)]).to(dev)
    out = torch.empty((16, a.shape[0], 100), device=dev)
    win = sharded.ExchangeWindow(frames=16, nb=100, list_cap=0)
    fn = lambda: sharded.anchor_assign_sharded(a, b, win, out=out)
elif what == "pib128":      # the bench's points_in_boxes call: 128 frames, per-frame boxes and points (SURVEY 8d generator)
    B = 128
    boxes = torch.stack([synth.waymo_boxes(200, 100 + f) for f in range(B)])
    pts = torch.stack([synth.points(180000, boxes[f], synth.WAYMO_RANGE, 0.05, seed=500 + f) for f in range(B)]).to(dev)
    boxes = boxes.to(dev)
    fn = lambda: R.points_in_boxes_gpu(pts, boxes)
elif what == "pib16":       # per-frame points (SURVEY 8d generator), 16 frames
    B = 16
    boxes = torch.stack([synth.waymo_boxes(200, 100 + f) for f in range(B)])
    pts = torch.stack([synth.points(180000, boxes[f], synth.WAYMO_RANGE, 0.05, seed=500 + f) for f in range(B)]).to(dev)
    boxes = boxes.to(dev)
    fn = lambda: R.points_in_boxes_gpu(pts, boxes)
else:
    raise SystemExit("unknown workload")

for _ in range(iters):
    fn()
torch.cuda.synchronize()
print("done", what)
