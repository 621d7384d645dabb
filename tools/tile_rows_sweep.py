"""Developer aid: sweep the row-tile height of iou_tile_kernel on the bench's 16-frame anchor sweep (needs the debug lib)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import synth
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "glenet_b200/lib/libglenet_geom_dbg.so"))
lib.glenet_boxes_iou_frames_gpu.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong,
                                            ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
dev = torch.device("cuda:0")
F = 16
a = synth.anchors_kitti3().to(dev)
b = torch.stack([synth.kitti_boxes(100, 100 + f + 1) for f in range(F)]).to(dev)
out = torch.empty((F, a.shape[0], 100), device=dev)


def run(frames):
    lib.glenet_boxes_iou_frames_gpu(1, a.data_ptr(), 0, a.shape[0], b.data_ptr(), 700, 100, out.data_ptr(), frames, None)


for frames in (16, 1):
    for tr in (0, 384, 352, 320, 288, 256, 224, 192, 160, 128, 96, 64):
        lib.glenet_debug_set_tile_rows(tr)
        for _ in range(3):
            run(frames)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            run(frames)
        e.record(); torch.cuda.synchronize()
        us = s.elapsed_time(e) / 10 * 1e3
        print(f"frames {frames:2d} TR {tr:3d} ({'auto' if tr == 0 else 'forced'}): {us:8.1f} us / launch, {us / frames:6.2f} us / frame")
