"""Developer aid: per-phase cycle breakdown, queue statistics and ablation timings of iou_tile_kernel.

    tools/build_debug_lib.sh && gpurun -- python tools/phase_timing.py
(needs the library built with -DGLENET_PHASE_TIMING)"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import synth
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "glenet_b200/lib/libglenet_geom_dbg.so"))
lib.glenet_boxes_iou_bev_gpu.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
dev = torch.device("cuda:0")
wl = sys.argv[1] if len(sys.argv) > 1 else "anchors"
if wl == "anchors":
    a = synth.anchors_kitti3().to(dev); b = synth.kitti_boxes(100, 4).to(dev)
elif wl == "dense":
    a = synth.proposals(1024, 20, 0)[0].to(dev); b = a
elif wl == "adversarial":
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    from make_golden import adversarial_boxes
    a = adversarial_boxes().to(dev); b = a
else:
    a = synth.kitti_boxes(6000, 0).to(dev); b = synth.kitti_boxes(100, 1).to(dev)
print("workload", wl, tuple(a.shape), tuple(b.shape), "resident CTAs", lib.glenet_debug_iou_resident_ctas())
outs = [torch.empty((a.shape[0], b.shape[0]), device=dev) for _ in range(4)]   # 4 x 84.5 MB > L2
buf = (ctypes.c_ulonglong * 12)()
names = ["0 stage+reduce", "1 active cols", "2 pair tests (+inner drains)", "3 compaction+prepare", "4 clip", "5 write results", "6 wait for fill/barrier", "7 sat filter"]


def run(out):
    lib.glenet_boxes_iou_bev_gpu(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], out.data_ptr(), None)


def timed(iters=20):
    for o in outs:
        run(o)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        run(outs[i % 4])
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


if len(sys.argv) > 2:
    lib.glenet_debug_set_col_split(int(sys.argv[2]))
    print("column split", sys.argv[2])
for flags, label in ((0, "full kernel"), (1, "no clip"), (2, "no zero fill"), (3, "neither"), (4, "STG fill"), (5, "STG fill, no clip")):
    lib.glenet_debug_set_flags(flags)
    print(f"{label:14s} {timed():7.2f} us / launch")
lib.glenet_debug_set_flags(int(os.environ.get('PHASE_FLAGS', '0')))
lib.glenet_debug_iou_phase_cycles(buf)
run(outs[0])
lib.glenet_debug_iou_phase_cycles(buf)
ncta = max(buf[11], 1)
tot = sum(buf[:8])
print(f"drains {buf[11]}  queued pairs/drain {buf[8] / ncta:.1f}  clipped pairs/drain {buf[9] / ncta:.1f}  boxes prepared/drain {buf[10] / ncta:.1f}")
print(f"cycles per CTA {tot / ncta:.0f}")
for n, v in zip(names, list(buf)[:8]):
    print(f"   {n:28s} {v / ncta:10.0f} cycles/CTA  {100 * v / max(tot, 1):5.1f}%")

import numpy as np
log = (ctypes.c_ulonglong * (4096 * 4))()
lib.glenet_debug_iou_cta_log(log, 4096)
L = np.array(log, dtype=np.uint64).reshape(4096, 4).astype(np.int64)
L = L[L[:, 0] > 0]
t0 = L[:, 0].min()
start, end, nq, nc = (L[:, 0] - t0) / 1e3, (L[:, 1] - t0) / 1e3, L[:, 2], L[:, 3]
dur = end - start
print(f"CTA start offset us: min {start.min():.1f} p50 {np.median(start):.1f} max {start.max():.1f};  end: p50 {np.median(end):.1f} p90 {np.percentile(end, 90):.1f} max {end.max():.1f}")
print(f"CTA duration us: min {dur.min():.1f} p50 {np.median(dur):.1f} p90 {np.percentile(dur, 90):.1f} max {dur.max():.1f}")
for lo, hi in ((0, 1), (1, 113), (113, 225), (225, 449), (449, 100000)):
    m = (nc >= lo) & (nc < hi)
    if m.any():
        print(f"   clipped pairs in [{lo},{hi}): {m.sum():4d} CTAs, duration p50 {np.median(dur[m]):.1f} max {dur[m].max():.1f} us, end max {end[m].max():.1f}")
