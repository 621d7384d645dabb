import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import iou3d_nms_utils as I, iou3d_utils as I1, synth
dev = torch.device('cuda:0')
s, g = synth.cvae_samples(20000, 30, 0); s, g = s.to(dev), g.to(dev)
pr, tg = synth.head_pairs(20000, 3); pr, tg = pr.to(dev), tg.to(dev)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e3
print('cvae aligned 600k: %.1f us' % t(lambda: I.boxes_iou3d_aligned(s, g, 30)))
print('v1 aligned 20k: %.1f us' % t(lambda: I1.boxes_aligned_iou3d_gpu(pr, tg)))
