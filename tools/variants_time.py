"""Developer aid: time the library variants built by tools/variants.sh through the C ABI (one GPU call for all of them)."""
import ctypes, glob, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glenet_b200 import synth
dev = torch.device("cuda:0")
vp = ctypes.c_void_p


def ev(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


smp, gt = synth.cvae_samples(20000, 30, 0); smp, gt = smp.to(dev), gt.to(dev)
al_out = torch.empty(600000, device=dev)
fb = torch.stack([synth.proposals(4096, 20, 20 + f)[0] for f in range(8)])
fs = torch.stack([synth.proposals(4096, 20, 20 + f)[1] for f in range(8)])
order = fs.sort(1, descending=True)[1]
nms_boxes = torch.gather(fb, 1, order.unsqueeze(-1).expand(-1, -1, 7)).contiguous().to(dev)
keep = torch.empty((8, 4096), dtype=torch.int64, device=dev); num = torch.zeros(8, dtype=torch.int32, device=dev)
g2 = synth.waymo_boxes(200, 2); pr = synth.proposals(4096, seed=3, base=g2)[0].to(dev); g2 = g2.to(dev)
d_out = torch.empty((4096, 200), device=dev)
anchors = synth.anchors_kitti3().to(dev)
gts = torch.stack([synth.kitti_boxes(100, 101 + f) for f in range(16)]).to(dev)
sweep_out = torch.empty((16, anchors.shape[0], 100), device=dev)
sq = synth.proposals(4096, 20, 7)[0].to(dev); sq_out = torch.empty((4096, 4096), device=dev)
ref = {}
for path in sorted(glob.glob(os.path.join(ROOT, "glenet_b200/lib/libglenet_geom_var_*.so"))):
    lib = ctypes.CDLL(path)
    lib.glenet_last_error.restype = ctypes.c_char_p
    lib.glenet_nms_workspace_bytes.restype = ctypes.c_size_t
    lib.glenet_boxes_iou_aligned_gpu.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, vp, vp]
    lib.glenet_nms_gpu.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_float, vp, vp, vp, ctypes.c_size_t, vp]
    lib.glenet_boxes_iou3d_gpu.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, vp, vp]
    lib.glenet_boxes_iou_bev_gpu.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, vp, vp]
    lib.glenet_boxes_iou_frames_gpu.argtypes = [ctypes.c_int, vp, ctypes.c_longlong, ctypes.c_int, vp, ctypes.c_longlong, ctypes.c_int, vp, ctypes.c_int, vp]
    ws_bytes = lib.glenet_nms_workspace_bytes(8, 4096)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    calls = {
        "cfg3_aligned_600k": lambda: lib.glenet_boxes_iou_aligned_gpu(2, smp.data_ptr(), 600000, gt.data_ptr(), 30, al_out.data_ptr(), None),
        "cfg1_nms_8x4096": lambda: lib.glenet_nms_gpu(nms_boxes.data_ptr(), 8, 4096, 0.7, keep.data_ptr(), num.data_ptr(), ws.data_ptr(), ws_bytes, None),
        "cfg2_iou3d_4096x200": lambda: lib.glenet_boxes_iou3d_gpu(pr.data_ptr(), 4096, g2.data_ptr(), 200, d_out.data_ptr(), None),
        "iou_bev_4096x4096": lambda: lib.glenet_boxes_iou_bev_gpu(sq.data_ptr(), 4096, sq.data_ptr(), 4096, sq_out.data_ptr(), None),
        "cfg4_sweep_16f": lambda: lib.glenet_boxes_iou_frames_gpu(1, anchors.data_ptr(), 0, anchors.shape[0], gts.data_ptr(), 700, 100, sweep_out.data_ptr(), 16, None),
        "cfg4_sweep_1f": lambda: lib.glenet_boxes_iou_frames_gpu(1, anchors.data_ptr(), 0, anchors.shape[0], gts.data_ptr(), 700, 100, sweep_out.data_ptr(), 1, None),
    }
    outs = {"cfg3_aligned_600k": al_out, "cfg1_nms_8x4096": num, "cfg2_iou3d_4096x200": d_out, "iou_bev_4096x4096": sq_out, "cfg4_sweep_16f": sweep_out}
    line = [f"{os.path.basename(path)[len('libglenet_geom_var_'):-3]:14s}"]
    for name, fn in calls.items():
        rc = fn(); torch.cuda.synchronize()
        if rc != 0:
            line.append(f"{name} ERROR {rc} {lib.glenet_last_error()}"); continue
        us = ev(fn)
        same = ""
        if name in outs:
            fn(); torch.cuda.synchronize()
            cur = outs[name].clone()
            if name not in ref:
                ref[name] = cur
            same = "=" if torch.equal(torch.nan_to_num(cur.float(), nan=-7.0), torch.nan_to_num(ref[name].float(), nan=-7.0)) else "DIFF"
        line.append(f"{name} {us:7.1f}us{same}")
    print("  ".join(line), flush=True)
