"""Developer aid: time the libpib_*.so variants built by tools/pib_variants.sh on BASELINE config 2
(128 frames x 180000 points x 200 boxes) and on a ragged single frame; every variant must return identical assignments."""
import ctypes, glob, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glenet_b200 import synth
dev = torch.device("cuda:0")
B, M, N = 128, 180000, 200
# the bench's input: every frame its own boxes and its own points, 5 % resampled inside that frame's boxes
boxes_h = torch.stack([synth.waymo_boxes(N, 100 + f) for f in range(B)])
pts = torch.stack([synth.points(M, boxes_h[f], synth.WAYMO_RANGE, 0.05, seed=500 + f) for f in range(B)]).to(dev).contiguous()
boxes = boxes_h.to(dev)
# second problem: ragged sizes (M not a multiple of anything, 3 frames, 77 boxes)
B2, M2, N2 = 3, 123457, 77
boxes2 = torch.stack([synth.waymo_boxes(N2, 7 + f) for f in range(B2)]).to(dev)
pts2 = torch.stack([synth.points(M2, boxes2[f].cpu(), synth.WAYMO_RANGE, 0.2, seed=11 + f) for f in range(B2)]).to(dev).contiguous()
# third problem: the reference's own call, one frame (launch / latency bound)
boxes3, pts3 = boxes[:1].contiguous(), pts[:1].contiguous()
ref = {}
for path in sorted(glob.glob(os.path.join(ROOT, "glenet_b200/lib/variants/libpib_*.so"))):
    lib = ctypes.CDLL(path)
    wsb = lib.glenet_points_in_boxes_workspace_bytes
    wsb.restype = ctypes.c_size_t; wsb.argtypes = [ctypes.c_int, ctypes.c_int]
    fn = lib.glenet_points_in_boxes_gpu
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    line = f"{os.path.basename(path):32s}"
    for tag, (bx, pt, b, n, m) in {"cfg2": (boxes, pts, B, N, M), "ragged": (boxes2, pts2, B2, N2, M2), "frame1": (boxes3, pts3, 1, N, M)}.items():
        nbytes = wsb(b, n)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.full((b, m), -7, dtype=torch.int32, device=dev)
        run = lambda: fn(bx.data_ptr(), pt.data_ptr(), b, n, m, out.data_ptr(), ws.data_ptr(), nbytes, None)
        rc = run(); torch.cuda.synchronize()
        if rc != 0:
            line += f" {tag}: error {rc}"; continue
        if tag not in ref:
            ref[tag] = out.clone()
        same = torch.equal(out, ref[tag])
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(20):
                run()
            e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e) / 20)
        line += f" {tag}: {best * 1e3:7.1f} us {b * m * 16 / best / 1e6:6.0f} GB/s same={same} |"
    print(line, flush=True)
