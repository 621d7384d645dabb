"""Generate the golden vectors in tests/golden/ by RUNNING THE REFERENCE.

    python tests/golden/make_golden.py cpu     # here: reference CPU functions, called through the
                                               # reference's own Python wrappers imported verbatim
                                               # from /root/reference -> tests/golden/cpu_golden.npz
    gpurun -- python tests/golden/make_golden.py gpu
                                               # on a B200: the reference's CUDA kernels (oracle/_ref)
                                               # -> gpurun_out/gpu_golden.npz (copy it to tests/golden/)
    python tests/golden/make_golden.py cpu_v1  # pcdet/ops/iou3d (boxes_aligned_iou3d_gpu's op), CPU function
                                               # -> tests/golden/cpu_golden_v1.npz
    gpurun -- python tests/golden/make_golden.py gpu_v1    # the same op's CUDA kernel through the reference's
                                               # wrapper imported verbatim -> gpurun_out/gpu_golden_v1.npz

The inputs are produced by the seeded generators of glenet_b200.synth plus a hand-written
adversarial set (identical boxes, shared edges, corners at 0.01 +- ulp from an edge, zero
padding boxes, huge headings).  The reference ships no tests or fixtures of its own, so these
files are what pins the oracle and the CUDA kernels.
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from glenet_b200 import synth  # noqa: E402
from oracle import ref as oref  # noqa: E402

REF = os.environ.get("GLENET_REFERENCE", "/root/reference")


def adversarial_boxes() -> torch.Tensor:
    base = torch.tensor([10.0, 5.0, -1.0, 3.9, 1.6, 1.5, 0.3])
    rows = [base.clone()]
    for dxy in (0.0, 1e-3, 0.00999, 0.01, 0.01001, 0.02, 0.5, 1.6, 1.61, 3.9, 3.91):
        for ang in (0.0, 0.3, 0.3 + math.pi / 2, 0.3 + math.pi, 1.57, -2.8):
            b = base.clone(); b[0] += dxy * math.cos(0.3); b[1] += dxy * math.sin(0.3); b[6] = ang; rows.append(b)
            b = base.clone(); b[0] -= dxy * math.sin(0.3); b[1] += dxy * math.cos(0.3); b[6] = ang; rows.append(b)
    rows += [torch.tensor([0.0, 0.0, 0.0, 2.0, 2.0, 2.0, 0.0]), torch.tensor([2.0, 0.0, 0.0, 2.0, 2.0, 2.0, 0.0]),
             torch.tensor([2.01, 0.0, 0.0, 2.0, 2.0, 2.0, 0.0]), torch.tensor([1.0, 1.0, 0.5, 2.0, 2.0, 2.0, math.pi / 4]),
             torch.zeros(7), torch.zeros(7), torch.tensor([0.0, 0.0, 0.0, 2.0, 2.0, 2.0, 1e4]),
             torch.tensor([0.5, 0.5, 0.0, 2.0, 1.0, 2.0, -1e4]), torch.tensor([70.0, 39.9, -1.0, 0.8, 0.6, 1.73, 1.57])]
    return torch.stack(rows).contiguous()


def inputs():
    d = {}
    d["sparse_a"], d["sparse_b"] = synth.kitti_boxes(96, 0), synth.kitti_boxes(40, 1)
    p, s = synth.proposals(160, 6, 0)
    d["dense"], d["dense_scores"] = p, s
    d["adv"] = adversarial_boxes()
    gt = synth.waymo_boxes(24, 2)
    pr, _ = synth.proposals(128, seed=3, base=gt)
    d["waymo_p"], d["waymo_gt"] = pr, gt
    nb, ns = synth.proposals(600, 12, 7)
    d["nms_boxes"], d["nms_scores"] = nb, ns
    pb = torch.stack([synth.kitti_boxes(16, 30), synth.kitti_boxes(16, 31)])
    pb[1, 10:] = 0  # zero padding rows as callers pass them
    pp = torch.stack([synth.points(6000, pb[0], synth.KITTI_RANGE, 0.3, seed=0), synth.points(6000, pb[1, :10], synth.KITTI_RANGE, 0.3, seed=1)])
    pp[0, 0] = pb[0, 0, :3]
    pp[0, 1] = pb[0, 0, :3] + torch.tensor([0.0, 0.0, 0.5]) * pb[0, 0, 5]
    pp[1, 0] = 0.0
    d["pib_boxes"], d["pib_points"] = pb, pp
    return d


def load_reference_wrappers():
    """Import the reference's wrapper modules verbatim, with stub parent packages."""
    def pkg(name):
        m = types.ModuleType(name); m.__path__ = []; sys.modules[name] = m; return m
    for n in ("pcdet", "pcdet.utils", "pcdet.ops", "pcdet.ops.iou3d_nms", "pcdet.ops.roiaware_pool3d"):
        pkg(n)
    sys.modules["SharedArray"] = types.ModuleType("SharedArray")
    sys.modules["pcdet.ops.iou3d_nms.iou3d_nms_cuda"] = oref.iou3d_nms_cuda()
    sys.modules["pcdet.ops.iou3d_nms"].iou3d_nms_cuda = oref.iou3d_nms_cuda()
    sys.modules["pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda"] = oref.roiaware_pool3d_cuda()
    sys.modules["pcdet.ops.roiaware_pool3d"].roiaware_pool3d_cuda = oref.roiaware_pool3d_cuda()

    def load(dotted, rel):
        spec = importlib.util.spec_from_file_location(dotted, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec); sys.modules[dotted] = m; spec.loader.exec_module(m); return m
    cu = load("pcdet.utils.common_utils", "pcdet/utils/common_utils.py")
    sys.modules["pcdet.utils"].common_utils = cu
    iou = load("pcdet.ops.iou3d_nms.iou3d_nms_utils", "pcdet/ops/iou3d_nms/iou3d_nms_utils.py")
    roi = load("pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils", "pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py")
    return iou, roi


def load_reference_iou3d_utils():
    """pcdet/ops/iou3d/iou3d_utils.py imported verbatim (its only native dependency is iou3d_cuda)."""
    def pkg(name):
        if name not in sys.modules:
            m = types.ModuleType(name); m.__path__ = []; sys.modules[name] = m
        return sys.modules[name]
    for n in ("pcdet", "pcdet.ops", "pcdet.ops.iou3d"):
        pkg(n)
    sys.modules["pcdet.ops.iou3d.iou3d_cuda"] = oref.iou3d_cuda()
    sys.modules["pcdet.ops.iou3d"].iou3d_cuda = oref.iou3d_cuda()
    spec = importlib.util.spec_from_file_location("pcdet.ops.iou3d.iou3d_utils", os.path.join(REF, "pcdet/ops/iou3d/iou3d_utils.py"))
    m = importlib.util.module_from_spec(spec); sys.modules["pcdet.ops.iou3d.iou3d_utils"] = m; spec.loader.exec_module(m)
    return m


def v1_inputs():
    pred, tgt = synth.head_pairs(600, 0)
    return {"v1_pred": pred, "v1_tgt": tgt}


def make_cpu_v1():
    """pcdet/ops/iou3d: the reference's own boxes3d_to_bev_torch + its CPU overlap (boxes_overlap_bev_cpu, diagonal)."""
    u = load_reference_iou3d_utils()
    d = v1_inputs()
    a5, b5 = u.boxes3d_to_bev_torch(d["v1_pred"]), u.boxes3d_to_bev_torch(d["v1_tgt"])
    full = torch.zeros((a5.shape[0], b5.shape[0]))
    oref.iou3d_cuda().boxes_overlap_bev_cpu(a5.contiguous(), b5.contiguous(), full)
    out = {k: v.numpy() for k, v in d.items()}
    out["v1_pred_bev"], out["v1_tgt_bev"] = a5.numpy(), b5.numpy()
    out["cpu_v1_overlap_aligned"] = full.diagonal().numpy().copy()
    out["cpu_v1_overlap_block"] = full[:64, :64].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "cpu_golden_v1.npz"), **out)
    print("wrote cpu_golden_v1.npz", {k: v.shape for k, v in out.items()})


def make_gpu_v1():
    """pcdet/ops/iou3d on a B200: the reference's Python wrapper verbatim around its compiled kernel."""
    # /root/reference does not exist on the GPU box: there the wrapper is oracle.ref's restatement of
    # iou3d_utils.py:332-387 (checked against the file itself wherever the reference tree is present)
    u = load_reference_iou3d_utils() if os.path.isdir(REF) else oref
    dev = torch.device("cuda:0")
    d = {k: v.to(dev) for k, v in v1_inputs().items()}
    iou3d, iou_bev = u.boxes_aligned_iou3d_gpu(d["v1_pred"], d["v1_tgt"], need_bev=True)
    iou3d_lwh = u.boxes_aligned_iou3d_gpu(d["v1_pred"], d["v1_tgt"], box_mode="lwh")
    out = {"gpu_v1_iou3d": iou3d.cpu().numpy(), "gpu_v1_iou_bev": iou_bev.cpu().numpy(), "gpu_v1_iou3d_lwh": iou3d_lwh.cpu().numpy()}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "gpu_golden_v1.npz"), **out)
    print("wrote gpurun_out/gpu_golden_v1.npz", {k: v.shape for k, v in out.items()})


def make_cpu():
    iou, roi = load_reference_wrappers()
    d = inputs()
    out = {k: v.numpy() for k, v in d.items()}
    out["cpu_iou_sparse"] = iou.boxes_bev_iou_cpu(d["sparse_a"], d["sparse_b"]).numpy()
    out["cpu_iou_dense"] = iou.boxes_bev_iou_cpu(d["dense"], d["dense"]).numpy()
    out["cpu_iou_adv"] = iou.boxes_bev_iou_cpu(d["adv"], d["adv"]).numpy()
    out["cpu_iou_waymo"] = iou.boxes_bev_iou_cpu(d["waymo_p"], d["waymo_gt"]).numpy()
    for f in range(2):
        out[f"cpu_pib_mask_{f}"] = np.packbits(roi.points_in_boxes_cpu(d["pib_points"][f], d["pib_boxes"][f]).numpy().astype(np.uint8), axis=1)
    # GLENet's variance-voting NMS (new_nms_gpu -> nms_func, iou3d_nms_utils.py:200-273), reference Python on CPU
    g = torch.Generator().manual_seed(11)
    vb, vs = synth.proposals(300, 10, 4)
    vvar = torch.rand((300, 7), generator=g) * 0.5 + 0.05
    out["vnms_boxes"], out["vnms_scores"], out["vnms_var"] = vb.numpy(), vs.numpy(), vvar.numpy()
    for name, kw in (("var", dict(variance=vvar.clone())), ("novar", dict()), ("thr", dict(variance=vvar.clone(), score_threshold=0.2))):
        keep, _, nb = iou.new_nms_gpu(vb.clone(), vs.clone(), 0.25, NMS_TYPE="new_nms_gpu", NMS_PRE_MAXSIZE=4096, **kw)
        out[f"vnms_keep_{name}"] = np.asarray(keep).astype(np.int64)
        out[f"vnms_newboxes_{name}"] = np.asarray(nb)[np.asarray(keep)]
    np.savez_compressed(os.path.join(HERE, "cpu_golden.npz"), **out)
    print("wrote cpu_golden.npz", {k: v.shape for k, v in out.items()})


def make_gpu():
    dev = torch.device("cuda:0")
    d = {k: v.to(dev) for k, v in inputs().items()}
    out = {}
    for name, a, b in (("sparse", d["sparse_a"], d["sparse_b"]), ("dense", d["dense"], d["dense"]),
                       ("adv", d["adv"], d["adv"]), ("waymo", d["waymo_p"], d["waymo_gt"])):
        out[f"gpu_iou_bev_{name}"] = oref.boxes_iou_bev(a, b).cpu().numpy()
        out[f"gpu_overlap_{name}"] = oref.boxes_overlap_bev(a, b).cpu().numpy()
        out[f"gpu_iou3d_{name}"] = oref.boxes_iou3d_gpu(a, b).cpu().numpy()
    for thr in (0.7, 0.1, 0.01):
        out[f"gpu_nms_{thr}"] = oref.nms_gpu(d["nms_boxes"], d["nms_scores"], thr)[0].cpu().numpy()
        out[f"gpu_nms_normal_{thr}"] = oref.nms_normal_gpu(d["nms_boxes"], d["nms_scores"], thr)[0].cpu().numpy()
    out["gpu_nms_pre100_0.7"] = oref.nms_gpu(d["nms_boxes"], d["nms_scores"], 0.7, pre_maxsize=100)[0].cpu().numpy()
    out["gpu_pib_index"] = oref.points_in_boxes_gpu(d["pib_points"], d["pib_boxes"]).cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "gpu_golden.npz"), **out)
    print("wrote gpurun_out/gpu_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "cpu"
    {"cpu": make_cpu, "gpu": make_gpu, "cpu_v1": make_cpu_v1, "gpu_v1": make_gpu_v1}[mode]()
