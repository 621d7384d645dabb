"""Developer aid: where the NMS sweep's time goes (needs tools/build_debug_lib.sh)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glenet_b200 import synth
lib = ctypes.CDLL(os.path.join(ROOT, "glenet_b200/lib/libglenet_geom_dbg.so"))
lib.glenet_nms_workspace_bytes.restype = ctypes.c_size_t
lib.glenet_nms_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
lib.glenet_nms_gpu.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
dev = torch.device("cuda:0")
b, s = synth.proposals(4096, 20, 21)
order = s.argsort(descending=True)
bs = b[order].contiguous().to(dev)
ws_bytes = lib.glenet_nms_workspace_bytes(1, 4096)
ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
keep = torch.empty((4096,), dtype=torch.int64, device=dev); num = torch.zeros((1,), dtype=torch.int32, device=dev)
buf = (ctypes.c_ulonglong * 4)()
for thr in (0.7, 0.1):
    for _ in range(3):
        lib.glenet_nms_gpu(bs.data_ptr(), 1, 4096, thr, keep.data_ptr(), num.data_ptr(), ws.data_ptr(), ws_bytes, None)
    lib.glenet_debug_sweep_cycles(buf)
    lib.glenet_nms_gpu(bs.data_ptr(), 1, 4096, thr, keep.data_ptr(), num.data_ptr(), ws.data_ptr(), ws_bytes, None)
    lib.glenet_debug_sweep_cycles(buf)
    steps = max(buf[3], 1)
    print(f"thresh {thr}: kept {int(num.item())}; per step: warp 0 busy {buf[0] / steps:.0f} cycles, job warp busy {buf[2] / steps:.0f}, whole step {buf[1] / steps:.0f}")
