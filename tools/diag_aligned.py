"""Diagnostic: where does boxes_iou3d_aligned differ from the reference block diagonal? (GPU box, needs oracle/_ref)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glenet_b200 import iou3d_nms_utils as I, synth
from oracle import ref
dev = torch.device("cuda:0")
smp, gt = synth.cvae_samples(20000, 30, 0)
smp, gt = smp.to(dev), gt.to(dev)
def blockdiag(fn):
    out = []
    for r0 in range(0, smp.shape[0], 6000):
        r1 = min(smp.shape[0], r0 + 6000); g0, g1 = r0 // 30, (r1 - 1) // 30 + 1
        full = fn(smp[r0:r1].contiguous(), gt[g0:g1].contiguous()); idx = torch.arange(r0, r1, device=dev)
        out.append(full[idx - r0, idx // 30 - g0])
    return torch.cat(out)
want = blockdiag(ref.boxes_iou_bev)
ours_pair = blockdiag(I.boxes_iou_bev)
got = I.boxes_iou_bev_aligned(smp, gt, 30)
for name, x in (("aligned kernel", got), ("tile kernel", ours_pair)):
    d = (x - want).abs()
    bad = torch.nonzero(d > 0).flatten()
    print(f"{name}: {bad.numel()} of {x.numel()} differ, max {float(d.max()):.3e}, > 1e-5: {int((d > 1e-5).sum())}")
    for i in bad[torch.argsort(d[bad], descending=True)][:12].tolist():
        print("   idx", i, "got", float(x[i]), "want", float(want[i]), "diff", float(d[i]))
        print("      a =", [float(v) for v in smp[i]], "b =", [float(v) for v in gt[i // 30]])
# repeatability
got2 = I.boxes_iou_bev_aligned(smp, gt, 30)
print("repeatable:", torch.equal(got, got2))
np.savez("gpurun_out/diag_aligned.npz", smp=smp.cpu().numpy(), gt=gt.cpu().numpy(), got=got.cpu().numpy(), want=want.cpu().numpy())
