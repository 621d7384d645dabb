"""Tuning aid: time the anchor sweep (16 frames and 1 frame) for forced tile heights (GLENET_IOU_TILE_ROWS), one subprocess each."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from glenet_b200 import iou3d_nms_utils as I, synth
dev = torch.device("cuda:0")
a = synth.anchors_kitti3().to(dev)
g = torch.stack([synth.kitti_boxes(100, 101 + f) for f in range(16)]).to(dev)
out = torch.empty((16, a.shape[0], 100), device=dev)
def ev(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3
t16 = ev(lambda: I.boxes_iou_bev_frames(a, g, out=out))
t1 = ev(lambda: I.boxes_iou_bev_frames(a, g[:1], out=out[:1]))
slab = a[:26432].contiguous()
t8 = ev(lambda: I.boxes_iou_bev_frames(slab, g, out=out[:, :26432].contiguous() if False else out.view(-1)[:16*26432*100].view(16, 26432, 100)))
print("16f %%.1f us  1f %%.1f us  world8-slab 16f %%.1f us" %% (t16, t1, t8))
''' % ROOT
for rows in sys.argv[1:] or ["0", "128", "192", "256", "320", "384"]:
    env = dict(os.environ)
    if rows != "0":
        env["GLENET_IOU_TILE_ROWS"] = rows
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(f"tile rows {rows:>4s}: {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.returncode}", flush=True)
