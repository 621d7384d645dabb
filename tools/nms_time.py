"""Developer aid: device time of the batched and the per-frame NMS on BASELINE config 1 (8 frames x 4096 proposals)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import iou3d_nms_utils as I, synth
dev = torch.device("cuda:0")
fb, fs = [], []
for f in range(8):
    b, s = synth.proposals(4096, 20, 20 + f); fb.append(b); fs.append(s)
fb, fs = torch.stack(fb).to(dev), torch.stack(fs).to(dev)
def ev(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3
for thr in (0.7, 0.1, 0.01):
    keep, num = I.nms_gpu_batch(fb, fs, thr)
    print("nms batch 8x4096 thr", thr, "%.1f us" % ev(lambda: I.nms_gpu_batch(fb, fs, thr)), "kept", num.tolist()[:4], "one frame %.1f us" % ev(lambda: I.nms_gpu(fb[0], fs[0], thr)))
