"""Developer check on a B200 box: new kernels vs the reference's own GPU kernels (oracle/_ref).

    gpurun -- python tools/gpu_check.py [--quick]

Prints parity statistics and device timings, writes gpurun_out/gpu_check.json.
(Development aid; the judged parity tests live in tests/.)
"""
from __future__ import annotations

import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import iou3d_nms_utils as I, roiaware_pool3d_utils as R, synth  # noqa: E402
from oracle import ref  # noqa: E402

OUT = {}
dev = torch.device("cuda:0")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def cmp(name, mine, refv):
    mine, refv = mine.float().cpu().numpy(), refv.float().cpu().numpy()
    d = np.abs(mine - refv)
    bad = np.isnan(d).sum()
    d = np.nan_to_num(d)
    res = dict(n=int(mine.size), exact=float((mine == refv).mean()), maxdiff=float(d.max()) if d.size else 0.0,
               n_gt_1e5=int((d > 1e-5).sum()), zero_mismatch=int(((mine == 0) != (refv == 0)).sum()), nan=int(bad),
               frac_pos=float((refv > 0).mean()) if d.size else 0.0)
    print(f"[parity] {name}: {res}", flush=True)
    OUT["parity_" + name] = res
    return res


def adversarial_boxes():
    base = torch.tensor([10.0, 5.0, -1.0, 3.9, 1.6, 1.5, 0.3])
    rows = [base.clone()]
    for dxy in (0.0, 1e-3, 0.00999, 0.01, 0.01001, 0.02, 0.5, 1.6, 1.61, 3.9, 3.91):
        for ang in (0.0, 0.3, 0.3 + math.pi / 2, 0.3 + math.pi, 1.57, -2.8):
            b = base.clone(); b[0] += dxy * math.cos(0.3); b[1] += dxy * math.sin(0.3); b[6] = ang; rows.append(b)
            b = base.clone(); b[0] -= dxy * math.sin(0.3); b[1] += dxy * math.cos(0.3); b[6] = ang; rows.append(b)
    # shared edges / axis aligned / zero padding / big headings
    rows += [torch.tensor([0.0, 0.0, 0.0, 2.0, 2.0, 2.0, 0.0]), torch.tensor([2.0, 0.0, 0.0, 2.0, 2.0, 2.0, 0.0]),
             torch.tensor([2.01, 0.0, 0.0, 2.0, 2.0, 2.0, 0.0]), torch.tensor([1.0, 1.0, 0.5, 2.0, 2.0, 2.0, math.pi / 4]),
             torch.zeros(7), torch.zeros(7), torch.tensor([0.0, 0.0, 0.0, 2.0, 2.0, 2.0, 1e4]),
             torch.tensor([0.5, 0.5, 0.0, 2.0, 1.0, 2.0, -1e4]), torch.tensor([70.0, 39.9, -1.0, 0.8, 0.6, 1.73, 1.57])]
    return torch.stack(rows).contiguous()


def iou_checks(quick):
    cases = {}
    a, b = synth.kitti_boxes(2000, 0), synth.kitti_boxes(300, 1)
    cases["sparse_kitti"] = (a, b)
    p, _ = synth.proposals(1024 if quick else 4096, 20, 0)
    cases["dense_self"] = (p, p)
    gt = synth.waymo_boxes(200, 2)
    pr, _ = synth.proposals(4096, seed=3, base=gt)
    cases["waymo_4096x200"] = (pr, gt)
    adv = adversarial_boxes()
    cases["adversarial"] = (adv, adv)
    anch = synth.anchors_kitti3()
    cases["anchors_x100"] = (anch if not quick else anch[:50000], synth.kitti_boxes(100, 4))
    far = synth.kitti_boxes(512, 5); far[:, 0] += 3000.0; far2 = synth.kitti_boxes(512, 5); far2[:, 0] += 3000.0
    far2[:, :2] += torch.randn(512, 2, generator=torch.Generator().manual_seed(1)) * 0.5
    cases["far_coords"] = (far, far2)
    for name, (a, b) in cases.items():
        a, b = a.to(dev), b.to(dev)
        cmp("iou_bev_" + name, I.boxes_iou_bev(a, b), ref.boxes_iou_bev(a, b))
        cmp("overlap_" + name, I.boxes_overlap_bev(a, b), ref.boxes_overlap_bev(a, b))
        cmp("iou3d_" + name, I.boxes_iou3d_gpu(a, b), ref.boxes_iou3d_gpu(a, b))
        t_new = timeit(lambda: I.boxes_iou_bev(a, b))
        t_ref = timeit(lambda: ref.boxes_iou_bev(a, b))
        t3_new = timeit(lambda: I.boxes_iou3d_gpu(a, b))
        t3_ref = timeit(lambda: ref.boxes_iou3d_gpu(a, b))
        OUT["time_iou_" + name] = dict(pairs=a.shape[0] * b.shape[0], new_ms=t_new, ref_ms=t_ref, new3d_ms=t3_new, ref3d_ms=t3_ref)
        print(f"[time] iou {name}: pairs={a.shape[0] * b.shape[0]} bev new {t_new:.4f} ms ref {t_ref:.4f} ms | 3d new {t3_new:.4f} ref {t3_ref:.4f}", flush=True)
    # aligned API vs block diagonal of the pairwise one
    s, g = synth.cvae_samples(2000, 30, 0)
    s, g = s.to(dev), g.to(dev)
    al = I.boxes_iou3d_aligned(s, g, 30)
    full = ref.boxes_iou3d_gpu(s[:3000], g[:100])
    diag = torch.stack([full[i, i // 30] for i in range(3000)])
    cmp("iou3d_aligned_cvae", al[:3000], diag)


def nms_checks(quick):
    for n in ((1000, 4096) if not quick else (1000,)):
        boxes, scores = synth.proposals(n, 20, 7)
        boxes, scores = boxes.to(dev), scores.to(dev)
        for thresh in (0.7, 0.1, 0.01, 0.85):
            for name, fn_new, fn_ref in (("nms", I.nms_gpu, ref.nms_gpu), ("nms_normal", I.nms_normal_gpu, ref.nms_normal_gpu)):
                k_new = fn_new(boxes, scores, thresh)[0]
                k_ref = fn_ref(boxes, scores, thresh)[0]
                same = bool(k_new.shape == k_ref.shape and torch.equal(k_new, k_ref))
                t_new = timeit(lambda: fn_new(boxes, scores, thresh), iters=5)
                t_ref = timeit(lambda: fn_ref(boxes, scores, thresh), iters=5)
                print(f"[parity] {name} n={n} thr={thresh}: equal={same} kept={k_ref.numel()} | new {t_new:.3f} ms ref {t_ref:.3f} ms", flush=True)
                OUT[f"{name}_{n}_{thresh}"] = dict(equal=same, kept=int(k_ref.numel()), new_ms=t_new, ref_ms=t_ref)
    # batched
    fb, fs = [], []
    for f in range(8):
        b, s = synth.proposals(4096, 20, 20 + f)
        fb.append(b); fs.append(s)
    fb, fs = torch.stack(fb).to(dev), torch.stack(fs).to(dev)
    keep, num = I.nms_gpu_batch(fb, fs, 0.7)
    ok = True
    for f in range(8):
        k_ref = ref.nms_gpu(fb[f], fs[f], 0.7)[0]
        ok &= bool(torch.equal(keep[f, :int(num[f])], k_ref))
    t_b = timeit(lambda: I.nms_gpu_batch(fb, fs, 0.7), iters=5)
    print(f"[parity] nms batch 8x4096: equal={ok} | {t_b:.3f} ms per batch", flush=True)
    OUT["nms_batch8"] = dict(equal=ok, ms=t_b)


def pib_checks(quick):
    for name, boxes_fn, rng, m, n, bsz in (("waymo", synth.waymo_boxes, synth.WAYMO_RANGE, 180000, 200, 4),
                                           ("kitti", synth.kitti_boxes, synth.KITTI_RANGE, 120000, 20, 3)):
        bs, ps = [], []
        for f in range(bsz):
            b = boxes_fn(n, 30 + f)
            if f == 1:
                b[n // 2:] = 0  # zero padding rows
            bs.append(b); ps.append(synth.points(m, b[: max(1, n // 2)], rng, 0.05, seed=f))
        boxes, pts = torch.stack(bs).to(dev), torch.stack(ps).to(dev)
        # adversarial points: exactly on faces / centre / z edge
        pts[0, 0] = boxes[0, 0, :3]
        pts[0, 1] = boxes[0, 0, :3] + torch.tensor([0., 0., 1.], device=dev) * boxes[0, 0, 5] / 2
        pts[1, 0] = 0
        mine = R.points_in_boxes_gpu(pts, boxes)
        refv = ref.points_in_boxes_gpu(pts, boxes)
        eq = bool(torch.equal(mine, refv))
        t_new = timeit(lambda: R.points_in_boxes_gpu(pts, boxes))
        t_ref = timeit(lambda: ref.points_in_boxes_gpu(pts, boxes))
        print(f"[parity] pib {name}: equal={eq} mismatches={(mine != refv).sum().item()} inside={(refv >= 0).sum().item()} | new {t_new:.4f} ms ref {t_ref:.4f} ms", flush=True)
        OUT["pib_" + name] = dict(equal=eq, mism=int((mine != refv).sum()), inside=int((refv >= 0).sum()), new_ms=t_new, ref_ms=t_ref, points=bsz * m)
    # overlapping boxes (first-hit order) + many boxes
    b, _ = synth.proposals(3000, 20, 9)
    pts = synth.points(60000, b[:20], synth.KITTI_RANGE, 0.5, seed=9)
    boxes, pts = b.unsqueeze(0).to(dev), pts.unsqueeze(0).to(dev)
    mine, refv = R.points_in_boxes_gpu(pts, boxes), ref.points_in_boxes_gpu(pts, boxes)
    print(f"[parity] pib overlapping 3000 boxes: equal={torch.equal(mine, refv)} inside={(refv >= 0).sum().item()}", flush=True)
    OUT["pib_overlap"] = dict(equal=bool(torch.equal(mine, refv)))
    # throughput config: B=128 x 180k x 200
    if not quick:
        B = 128
        boxes = torch.stack([synth.waymo_boxes(200, 100 + f) for f in range(B)]).to(dev)
        base = synth.points(180000, boxes[0].cpu(), synth.WAYMO_RANGE, 0.05, seed=5).to(dev)
        pts = base.unsqueeze(0).repeat(B, 1, 1).contiguous()
        pts += torch.randn_like(pts) * 0.01
        t_new = timeit(lambda: R.points_in_boxes_gpu(pts, boxes), iters=5)
        mine = R.points_in_boxes_gpu(pts, boxes)
        refv = ref.points_in_boxes_gpu(pts, boxes)
        t_ref = timeit(lambda: ref.points_in_boxes_gpu(pts, boxes), iters=2, warm=1)
        npts = B * 180000
        print(f"[time] pib 128x180k x200: equal={torch.equal(mine, refv)} new {t_new:.3f} ms = {npts / t_new / 1e6:.1f} Gpts/s ({npts * 16 / t_new / 1e6:.0f} GB/s) | ref {t_ref:.3f} ms", flush=True)
        OUT["pib_big"] = dict(equal=bool(torch.equal(mine, refv)), new_ms=t_new, ref_ms=t_ref, gpts=npts / t_new / 1e6)


def cpu_dialect_checks():
    a, b = synth.kitti_boxes(200, 0), synth.kitti_boxes(50, 1)
    t0 = time.perf_counter(); r = ref.boxes_bev_iou_cpu(a, b); t_ref = time.perf_counter() - t0
    t0 = time.perf_counter(); m = I.boxes_bev_iou_cpu(a, b); t_new = time.perf_counter() - t0
    cmp("cpu_iou_200x50", m, r)
    p, _ = synth.proposals(1500, 20, 0)
    t0 = time.perf_counter(); r = ref.boxes_bev_iou_cpu(p, p); t_ref2 = time.perf_counter() - t0
    t0 = time.perf_counter(); m = I.boxes_bev_iou_cpu(p, p); t_new2 = time.perf_counter() - t0
    cmp("cpu_iou_dense1500", m, r)
    adv = adversarial_boxes()
    cmp("cpu_iou_adv", I.boxes_bev_iou_cpu(adv, adv), ref.boxes_bev_iou_cpu(adv, adv))
    bx = synth.kitti_boxes(20, 3)
    pts = synth.points(120000, bx, synth.KITTI_RANGE, 0.05, seed=3)
    t0 = time.perf_counter(); r = ref.points_in_boxes_cpu(pts, bx); t_ref3 = time.perf_counter() - t0
    m = R.points_in_boxes_cpu(pts, bx)
    t0 = time.perf_counter(); m = R.points_in_boxes_cpu(pts, bx); t_new3 = time.perf_counter() - t0
    print(f"[parity] pib_cpu 120k x 20: equal={torch.equal(m, r)} inside={int(r.sum())} | new {t_new3 * 1e3:.2f} ms ref {t_ref3 * 1e3:.2f} ms", flush=True)
    print(f"[time] cpu iou 200x50 new {t_new * 1e3:.2f} ms ref {t_ref * 1e3:.2f} ms ; dense 1500^2 new {t_new2 * 1e3:.2f} ms ref {t_ref2 * 1e3:.2f} ms", flush=True)
    OUT["cpu_dialect"] = dict(pib_equal=bool(torch.equal(m, r)), iou_small_new_ms=t_new * 1e3, iou_small_ref_ms=t_ref * 1e3,
                              iou_dense_new_ms=t_new2 * 1e3, iou_dense_ref_ms=t_ref2 * 1e3, pib_new_ms=t_new3 * 1e3, pib_ref_ms=t_ref3 * 1e3)


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    print(torch.cuda.get_device_name(0), "cpus", os.cpu_count(), flush=True)
    for fn in (pib_checks, iou_checks, nms_checks):
        try:
            fn(quick)
        except Exception as ex:  # keep going: one failure must not hide the other results
            import traceback; traceback.print_exc()
            OUT[fn.__name__ + "_error"] = repr(ex)
    try:
        cpu_dialect_checks()
    except Exception as ex:
        import traceback; traceback.print_exc()
        OUT["cpu_dialect_error"] = repr(ex)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gpu_check.json", "w") as f:
        json.dump(OUT, f, indent=1)
