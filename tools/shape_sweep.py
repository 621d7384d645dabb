"""Developer aid: time the CTA-shape variants built by tools/shape_sweep.sh on the bench's anchor sweep (16 frames and 1 frame)."""
import ctypes, glob, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glenet_b200 import synth
dev = torch.device("cuda:0")
F = 16
a = synth.anchors_kitti3().to(dev)
b = torch.stack([synth.kitti_boxes(100, 100 + f + 1) for f in range(F)]).to(dev)
out = torch.empty((F, a.shape[0], 100), device=dev)
ref_out = None
for path in sorted(glob.glob(os.path.join(ROOT, "glenet_b200/lib/libglenet_geom_shape_*.so"))):
    lib = ctypes.CDLL(path)
    fn = lib.glenet_boxes_iou_frames_gpu
    fn.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    res = {}
    for frames in (16, 1):
        run = lambda: fn(1, a.data_ptr(), 0, a.shape[0], b.data_ptr(), 700, 100, out.data_ptr(), frames, None)
        rc = run()
        torch.cuda.synchronize()
        if rc != 0:
            lib.glenet_last_error.restype = ctypes.c_char_p
            print(os.path.basename(path), "error", rc, lib.glenet_last_error()); break
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            run()
        e.record(); torch.cuda.synchronize()
        res[frames] = s.elapsed_time(e) / 20 * 1e3
    if len(res) == 2:
        fn(1, a.data_ptr(), 0, a.shape[0], b.data_ptr(), 700, 100, out.data_ptr(), 16, None)
        torch.cuda.synchronize()
        if ref_out is None:
            ref_out = out.clone()
        same = torch.equal(out, ref_out)
        print(f"{os.path.basename(path):40s} 16 frames {res[16]:7.1f} us ({res[16] / 16:5.2f} us/frame)   1 frame {res[1]:6.1f} us   identical to base: {same}")
