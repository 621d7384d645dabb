"""Developer aid: wall clock of GLENet's variance-voting NMS (new_nms_gpu) on 4096 proposals."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import iou3d_nms_utils as I, synth
dev = torch.device("cuda:0")
p, s = synth.proposals(4096, 20, 3)
var = (torch.rand((4096, 7), generator=torch.Generator().manual_seed(1)) * 0.5 + 0.05)
p, s, var = p.to(dev), s.to(dev), var.to(dev)
for _ in range(3): I.new_nms_gpu(p, s, 0.25, variance=var)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): keep, _, voted = I.new_nms_gpu(p, s, 0.25, variance=var)
torch.cuda.synchronize(); print("new_nms_gpu 4096: %.3f ms per frame, kept %d" % ((time.perf_counter() - t0) / 10 * 1e3, len(keep)))
for _ in range(3): I.softnms_gpu(p, s, 0.25, score_threshold=0.1, soft_mode="gaussian", soft_sigma=0.3, variance=var)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): I.softnms_gpu(p, s, 0.25, score_threshold=0.1, soft_mode="gaussian", soft_sigma=0.3, variance=var)
torch.cuda.synchronize(); print("softnms_gpu 4096: %.3f ms per frame" % ((time.perf_counter() - t0) / 10 * 1e3))
