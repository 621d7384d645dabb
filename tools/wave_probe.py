"""Developer aid: does the anchor sweep time jump when the tile count crosses one wave (592 CTAs)?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glenet_b200 import iou3d_nms_utils as I, synth
dev = torch.device("cuda:0")
anchors = synth.anchors_kitti3().to(dev)
gts = [synth.kitti_boxes(100, 100 + f).to(dev) for f in range(16)]
for na in (256 * 148, 256 * 296, 256 * 444, 256 * 592, 256 * 593, 256 * 650, 256 * 740, 211200):
    a = anchors[:na].contiguous()
    def run():
        for f in range(16): I.boxes_iou_bev(a, gts[f])
    for _ in range(3): run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): run()
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) / 160 * 1000
    print(f"na={na:7d} tiles={(na + 255) // 256:4d}: {us:6.1f} us per launch, {na * 100 / us / 1e3:7.1f} Gpairs/s", flush=True)
