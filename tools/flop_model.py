"""Algorithmic FLOPs of the benchmark's seeded inputs ("the oracle restatement exports (k1, k2, c, cnt) histograms
for each seeded input so that sum F is a constant of the benchmark, not of the implementation").

    python tools/flop_model.py            # -> profiles/flop_model.json   (CPU only, ~1 min; uses oracle/geom_oracle.c)

F_pair = 8 for a pair the exact circle test culls, else 64 + 80 + 32 k1 + 19 k2 + 2 c + [cnt > 0] (27 cnt + cnt (cnt - 1) / 2
+ 8 (cnt - 1)) + 7 (add / sub / mul / compare = 1, FMA = 2, div = 1, atan2f = 25; per-box work hoisted).  bench.py reads the
JSON; it never executes the oracle on the measured path.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glenet_b200 import synth  # noqa: E402
from oracle import capi        # noqa: E402

F_CULL, F_BASE = 8, 64 + 80 + 7


def stats_of(a, b):
    _, st = capi.boxes_iou_bev(a, b, dialect=capi.GPU, stats=True)
    passed = capi.count_circle_pass(a, b)
    return {"pairs": int(st.pairs), "circle_pass": int(passed), "flops_all_heavy": st.flops(), "k1": int(st.k1), "k2": int(st.k2),
            "corners": int(st.corners), "cnt_hist": [int(x) for x in st.cnt_hist]}


def model_flops(s):
    """every pair evaluated by the oracle is counted on the heavy path; culled ones cost F_CULL instead of the base terms"""
    culled = s["pairs"] - s["circle_pass"]
    return s["flops_all_heavy"] - culled * (F_BASE - F_CULL)


def add(acc, s):
    for k, v in s.items():
        if isinstance(v, list):
            acc[k] = [x + y for x, y in zip(acc.get(k, [0] * len(v)), v)]
        else:
            acc[k] = acc.get(k, 0) + v
    return acc


def main():
    capi.load()
    out = {"model": "SURVEY.md 8d F_pair", "f_cull": F_CULL}
    # cfg3: 600 000 aligned pairs, bench seed 0: sample j of GT g against GT g
    smp, gt = synth.cvae_samples(20000, 30, 0)
    acc = {}
    for g in range(gt.shape[0]):
        add(acc, stats_of(smp[g * 30:(g + 1) * 30], gt[g:g + 1]))
    acc["flops"] = model_flops(acc)
    acc["flops_per_pair"] = acc["flops"] / acc["pairs"]
    out["cfg3_aligned_600k"] = acc
    # cfg1: 8 frames x upper triangle of 4096 proposals (bench seeds 20..27), boxes in score order does not matter for the sum
    tot = {}
    for f in range(8):
        boxes, _ = synth.proposals(4096, 20, 20 + f)
        full = stats_of(boxes, boxes)
        diag = {}
        for i in range(0, 4096, 1):
            add(diag, stats_of(boxes[i:i + 1], boxes[i:i + 1]))
        tri = {}
        for k in full:
            if isinstance(full[k], list):
                tri[k] = [(x - y) // 2 for x, y in zip(full[k], diag[k])]
            elif k == "flops_all_heavy":
                tri[k] = (full[k] - diag[k]) / 2
            else:
                tri[k] = (full[k] - diag[k]) // 2
        add(tot, tri)
    tot["flops"] = model_flops(tot)
    tot["flops_per_frame"] = tot["flops"] / 8
    out["cfg1_nms_8x4096_upper_triangle"] = tot
    # cfg2: boxes_iou3d_gpu 4096 x 200 (bench seeds)
    g2 = synth.waymo_boxes(200, 2)
    pr, _ = synth.proposals(4096, seed=3, base=g2)
    s = stats_of(pr, g2)
    s["flops"] = model_flops(s)
    out["cfg2_iou3d_4096x200"] = s
    # cfg4: one frame of the anchor sweep (frame seed 101) -- which roofline binds: FLOPs per pair vs 4 B per pair
    s = stats_of(synth.anchors_kitti3(), synth.kitti_boxes(100, 101))
    s["flops"] = model_flops(s)
    s["flops_per_pair"] = s["flops"] / s["pairs"]
    out["cfg4_frame_211200x100"] = s
    path = os.path.join(ROOT, "profiles", "flop_model.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path, {k: (v.get("flops"), v.get("pairs")) for k, v in out.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()
