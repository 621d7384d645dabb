"""Developer aid: per-phase cycle breakdown of iou_tile_kernel (needs lib built with -DGLENET_PHASE_TIMING)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import synth
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "glenet_b200/lib/libglenet_geom_dbg.so"))
lib.glenet_boxes_iou_bev_gpu.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
dev = torch.device("cuda:0")
a = synth.anchors_kitti3().to(dev); b = synth.kitti_boxes(100, 4).to(dev)
out = torch.empty((a.shape[0], b.shape[0]), device=dev)
buf = (ctypes.c_ulonglong * 8)()
names = ["0 stage+reduce", "1 active cols", "2 pair tests", "3 lazy prepare", "4 zero fill tail+barrier", "5 write results", "6 (unused)", "7 clip"]
for it in range(3):
    lib.glenet_debug_iou_phase_cycles(buf)
    lib.glenet_boxes_iou_bev_gpu(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], out.data_ptr(), None)
    lib.glenet_debug_iou_phase_cycles(buf)
    tot = sum(buf)
    print(f"iter {it}: total CTA cycles (sum over 825 CTAs) {tot}, per CTA {tot / 825:.0f}")
    for n, v in zip(names, buf):
        print(f"   {n:28s} {v / 825:10.0f} cycles/CTA  {100 * v / max(tot, 1):5.1f}%")
