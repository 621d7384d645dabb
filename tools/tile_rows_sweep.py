"""Developer aid: time the anchor sweep for several row-tile heights (debug library)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glenet_b200 import synth
lib = ctypes.CDLL(os.path.join(ROOT, "glenet_b200/lib/libglenet_geom_dbg.so"))
lib.glenet_boxes_iou_bev_gpu.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
dev = torch.device("cuda:0")
a = synth.anchors_kitti3().to(dev)
gts = [synth.kitti_boxes(100, 100 + f).to(dev) for f in range(16)]
outs = [torch.empty((a.shape[0], 100), device=dev) for _ in range(16)]
for tr, fl in ((0, 0), (0, 1), (0, 2), (0, 3), (128, 1), (128, 2)):
    lib.glenet_debug_set_tile_rows(tr); lib.glenet_debug_set_flags(fl)
    def run():
        for f in range(16):
            lib.glenet_boxes_iou_bev_gpu(a.data_ptr(), a.shape[0], gts[f].data_ptr(), 100, outs[f].data_ptr(), None)
    for _ in range(3): run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): run()
    
    e.record(); torch.cuda.synchronize()
    print(f"TR={tr:4d} flags={fl} (1=no clip, 2=no zero fill): {s.elapsed_time(e) / 160 * 1000:.1f} us per launch", flush=True)
# reference points: plain memset of one output matrix, and 16 of them
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    for o in outs: o.zero_()
s.record()
for _ in range(10):
    for o in outs: o.zero_()
e.record(); torch.cuda.synchronize()
print(f"torch zero_() of one (211200,100) f32 matrix: {s.elapsed_time(e) / 160 * 1000:.1f} us", flush=True)
