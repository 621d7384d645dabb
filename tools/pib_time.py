"""Developer aid: device time of points_in_boxes_gpu on BASELINE config 2 (128 frames x 180000 points x 200 boxes)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import roiaware_pool3d_utils as R, synth
dev = torch.device("cuda:0")
B, M = 128, 180000
boxes = torch.stack([synth.waymo_boxes(200, 100 + f) for f in range(B)]).to(dev)
base = synth.points(M, boxes[0].cpu(), synth.WAYMO_RANGE, 0.05, seed=5).to(dev)
pts = (base.unsqueeze(0).repeat(B, 1, 1) + torch.randn(B, M, 3, device=dev) * 0.01).contiguous()
for _ in range(3):
    R.points_in_boxes_gpu(pts, boxes)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        R.points_in_boxes_gpu(pts, boxes)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    best = min(best, ms)
    print(f"points_in_boxes_gpu cfg2: {ms * 1e3:.1f} us / call, {B * M / ms / 1e6:.1f} Gpts/s, {B * M * 16 / ms / 1e6:.0f} GB/s algorithmic")
