"""Small run of every kernel for compute-sanitizer (memcheck / racecheck):  compute-sanitizer --tool memcheck python tools/sanitize_workload.py"""
import os, sys
os.environ.setdefault("GLENET_NMS_SPATIAL", "2")   # the spatial-tile NMS kernels are taken by size in production: force them on the small inputs here
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import iou3d_nms_utils as I, iou3d_utils as I1, roiaware_pool3d_utils as R, synth
dev = torch.device("cuda:0")
a = synth.anchors_kitti3()[:3000].to(dev)
g = torch.stack([synth.kitti_boxes(37, 5 + f) for f in range(3)]).to(dev)
I.boxes_iou_bev_frames(a, g); I.boxes_iou3d_gpu_frames(a, g); I.boxes_overlap_bev(a, g[0])
g4 = torch.stack([synth.kitti_boxes(100, 9 + f) for f in range(2)]).to(dev)
I.boxes_iou_bev_frames(a, g4)                                   # bulk-copy zero fill path
p, s = synth.proposals(700, 6, 1)
p, s = p.to(dev), s.to(dev)
I.boxes_iou_bev(p, p)                                           # dense tiles: queue overflow rounds, carried clip passes
I.boxes_iou_frames_sparse(a, g4); I.iou_max_overlaps_frames(a, g4)
I.nms_gpu(p, s, 0.7); I.nms_gpu(p[:300], s[:300], -1.0)         # spatial tiles (forced above); a negative threshold takes the score-order tiles and the in-tile clip
I.nms_normal_gpu(p, s, 0.7); I.nms_gpu_batch(torch.stack([p, p]), torch.stack([s, s]), 0.5)
I.boxes_iou3d_aligned(p[:600], p[:20], 30)
pr, tg = synth.head_pairs(500, 0)
I1.boxes_aligned_iou3d_gpu(pr.to(dev), tg.to(dev), need_bev=True)
bx = torch.stack([synth.waymo_boxes(40, 30 + f) for f in range(2)])
pts = torch.stack([synth.points(9000, bx[f], synth.WAYMO_RANGE, 0.1, seed=f) for f in range(2)])
R.points_in_boxes_gpu(pts.to(dev), bx.to(dev)); R.points_in_boxes_gpu(pts.to(dev), bx[:, :5].contiguous().to(dev))
pk = synth.points(6000, p[:20].cpu(), synth.KITTI_RANGE, 0.5, seed=3)[None].to(dev)
R.points_in_boxes_gpu(pk, p[None, :300].contiguous())           # > 224 boxes: records through L1 / L2; crowded cells: list slices
R.points_in_boxes_cpu(pts[0], bx[0]); I.boxes_bev_iou_cpu(p[:50].cpu(), p[:40].cpu())
# variance-voting / soft NMS (vnms.cu), the single-rank exchange window (assign + gather + scatter + decode kernels)
v = (torch.rand((700, 7), generator=torch.Generator().manual_seed(5)) * 0.5 + 0.05).to(dev)
I.new_nms_gpu(p, s, 0.25, variance=v); I.softnms_gpu(p[:200], s[:200], 0.25, score_threshold=0.1, soft_mode="gaussian", soft_sigma=0.3, variance=v[:200])
from glenet_b200 import sharded
win = sharded.ExchangeWindow(frames=2, nb=100, list_cap=1 << 16)
sharded.anchor_assign_sharded(a, g4, win); sharded.anchor_assign_sharded(a, g4, win, dense=False); sharded.boxes_iou_gather_sharded(a, g4, win)
assert win.status() == 0
win.close()
# next scope rows: KITTI-evaluator rotated IoU (dense + blocks), GT-database crops (both rules), CVAE recall IoU
import numpy as np
from glenet_b200 import cvae_eval_utils as C, gt_database as G, rotate_iou as RI
rng = np.random.default_rng(0)
rb = np.concatenate([rng.uniform(0, 30, (150, 2)), rng.uniform(1.5, 4.3, (150, 2)), rng.uniform(-3.2, 3.2, (150, 1))], 1).astype(np.float32)
RI.rotate_iou_gpu_eval(rb, rb[:131], -1); RI.rotate_iou_gpu_eval_blocks(rb, rb[:131], [70, 0, 80], [60, 1, 70], 2)
G.crop_gt_objects(torch.cat([pts[0], torch.rand(9000, 1)], 1).numpy(), bx[0].numpy().astype(np.float64), "kitti")
G.crop_gt_objects(torch.cat([pts[0], torch.rand(9000, 2)], 1).numpy(), bx[0].numpy(), "waymo")
smp, gt = synth.cvae_samples(300, 1, 2)
C.iou3d(gt.to(dev), smp.reshape(-1, 7).to(dev))
torch.cuda.synchronize()
print("sanitize workload done")
