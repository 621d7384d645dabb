import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import iou3d_nms_utils as I, synth
from oracle import ref
dev = torch.device("cuda:0")
p, _ = synth.proposals(4096, 20, 0)
p = p.to(dev)
mine, want = I.boxes_iou_bev(p, p), ref.boxes_iou_bev(p, p)
bad = (mine != want).nonzero()
print("mismatches", bad.shape[0], "on diagonal", int((bad[:, 0] == bad[:, 1]).sum()))
for r, c in bad[:8].tolist():
    one = I.boxes_iou_bev(p[r:r + 1], p[c:c + 1])
    al = I.boxes_iou_bev_aligned(p[r:r + 1], p[c:c + 1], 1)
    sub = I.boxes_iou_bev(p[r - r % 8: r - r % 8 + 8], p[c - c % 4: c - c % 4 + 4])
    print(r, c, "full", float(mine[r, c]), "ref", float(want[r, c]), "1x1", float(one), "aligned", float(al), "8x4", float(sub[r % 8, c % 4]))
# repeatability
mine2 = I.boxes_iou_bev(p, p)
print("repeat equal", torch.equal(mine, mine2), "mismatch now", int((mine2 != want).sum()))
