"""Host-side logic that needs no GPU: shim registration, wrapper pre-launch checks, synthetic inputs."""
import importlib
import sys

import numpy as np
import pytest
import torch

from glenet_b200 import iou3d_nms_utils as I
from glenet_b200 import iou3d_utils as I1
from glenet_b200 import roiaware_pool3d_utils as R
from glenet_b200 import shim, synth


def _drop_pcdet_modules():
    for name in [k for k in sys.modules if k == "pcdet" or k.startswith("pcdet.")]:
        del sys.modules[name]


def test_shim_registers_reference_module_paths():
    _drop_pcdet_modules()
    patched = shim.install()
    m = importlib.import_module("pcdet.ops.iou3d_nms.iou3d_nms_utils")
    assert m.nms_gpu is I.nms_gpu and m.boxes_iou3d_gpu is I.boxes_iou3d_gpu and m.new_nms_gpu is I.new_nms_gpu
    from pcdet.ops.roiaware_pool3d import roiaware_pool3d_utils as r2
    assert r2.points_in_boxes_gpu is R.points_in_boxes_gpu
    # without pcdet installed the host-signature functions are provided too
    assert r2.points_in_boxes_cpu is R.points_in_boxes_cpu and m.boxes_bev_iou_cpu is I.boxes_bev_iou_cpu
    assert "nms_gpu" in patched["pcdet.ops.iou3d_nms.iou3d_nms_utils"]
    # the plugin mechanism of the reference: lookup by config string (model_nms_utils.py:40-52)
    for name in ("nms_gpu", "nms_normal_gpu"):
        assert callable(getattr(sys.modules["pcdet.ops.iou3d_nms.iou3d_nms_utils"], name))
    for name in ("boxes_bev_iou_cpu", "boxes_iou_bev", "boxes_iou3d_gpu", "points_in_boxes_cpu", "points_in_boxes_gpu"):
        assert callable(getattr(I, name, None) or getattr(R, name))


def test_wrappers_reject_bad_input_before_any_launch():
    a = synth.kitti_boxes(3, 0)
    with pytest.raises(AssertionError):
        I.boxes_iou_bev(a[:, :6], a)
    with pytest.raises(AssertionError):
        I.boxes_iou3d_gpu(a, a[:, :5])
    with pytest.raises(RuntimeError):
        I.boxes_iou_bev(a, a)                 # CPU tensors into the GPU entry point: raise, never fall back
    with pytest.raises(RuntimeError):
        R.points_in_boxes_gpu(torch.zeros(1, 4, 3), torch.zeros(1, 2, 7))
    with pytest.raises(AssertionError):
        R.points_in_boxes_cpu(np.zeros((4, 2), np.float32), np.zeros((2, 7), np.float32))
    with pytest.raises(AssertionError):
        R.points_in_boxes_gpu(torch.zeros(1, 4, 3), torch.zeros(2, 2, 7))


def test_additive_apis_reject_bad_input_before_any_launch():
    a = synth.kitti_boxes(6, 0)
    g = torch.stack([synth.kitti_boxes(4, 1), synth.kitti_boxes(4, 2)])
    for fn in (I.boxes_iou_bev_frames, I.boxes_iou3d_gpu_frames, I.boxes_overlap_bev_frames, I.iou_max_overlaps_frames, I.boxes_iou_frames_sparse):
        with pytest.raises(RuntimeError):
            fn(a, g)                           # CPU tensors: raise, never compute on the host
    pred, tgt = synth.head_pairs(5, 0)
    with pytest.raises(NotImplementedError):
        I1.boxes_aligned_iou3d_gpu(pred, tgt, rect=True)        # as the reference (iou3d_utils.py:356-357)
    with pytest.raises(AssertionError):
        I1.boxes_aligned_iou3d_gpu(pred, tgt[:3])
    with pytest.raises(RuntimeError):
        I1.boxes_aligned_iou3d_gpu(pred, tgt)
    with pytest.raises(ValueError):
        I1.boxes_aligned_iou3d_gpu(pred, tgt, box_mode="xyz")   # str.index, as in the reference
    with pytest.raises(AssertionError):
        I1.boxes_aligned_overlap_bev_cpu(pred[:, :5].cuda() if torch.cuda.is_available() else pred[:, :4], tgt[:, :5])
    bev = I1.boxes3d_to_bev_torch(pred, "lwh")
    assert bev.shape == (5, 5) and torch.equal(bev[:, 4], pred[:, 6])
    assert torch.equal(bev[:, 2] - bev[:, 0], (pred[:, 0] + pred[:, 4] / 2) - (pred[:, 0] - pred[:, 4] / 2))   # 'lwh': w = column 4


def test_shim_patches_a_real_checkout_function_by_function(tmp_path, monkeypatch):
    """Inside a real pcdet checkout the original modules stay and keep everything this package does not provide
    (RoIAwarePool3d for pcdet/models/roi_heads/partA2_head.py, the other functions of pcdet.ops.iou3d.iou3d_utils);
    the _cpu entry points keep the reference's host code unless asked for (forked DataLoader workers)."""
    _drop_pcdet_modules()
    root = tmp_path / "checkout"
    for pkg in ("pcdet", "pcdet/ops", "pcdet/ops/iou3d_nms", "pcdet/ops/roiaware_pool3d", "pcdet/ops/iou3d"):
        (root / pkg).mkdir(parents=True, exist_ok=True)
        (root / pkg / "__init__.py").write_text("")
    (root / "pcdet/ops/iou3d_nms/iou3d_nms_utils.py").write_text(
        "def nms_gpu(*a, **k):\n    return 'ref'\ndef boxes_bev_iou_cpu(*a):\n    return 'ref-cpu'\ndef some_other_helper():\n    return 7\n")
    (root / "pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py").write_text(
        "class RoIAwarePool3d:\n    pass\ndef points_in_boxes_cpu(*a):\n    return 'ref-cpu'\ndef points_in_boxes_gpu(*a):\n    return 'ref'\n")
    (root / "pcdet/ops/iou3d/iou3d_utils.py").write_text(
        "def nms_gpu(*a):\n    return 'ref-v1'\ndef boxes_iou3d_gpu(*a):\n    return 'ref-v1'\ndef boxes_aligned_iou3d_gpu(*a):\n    return 'ref'\n")
    monkeypatch.syspath_prepend(str(root))
    importlib.invalidate_caches()
    try:
        shim.install()
        m = importlib.import_module("pcdet.ops.iou3d_nms.iou3d_nms_utils")
        r = importlib.import_module("pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils")
        v1 = importlib.import_module("pcdet.ops.iou3d.iou3d_utils")
        assert m.__file__.startswith(str(root)) and m.nms_gpu is I.nms_gpu and m.some_other_helper() == 7
        assert m.boxes_bev_iou_cpu() == "ref-cpu" and r.points_in_boxes_cpu() == "ref-cpu"        # left alone by default
        assert hasattr(r, "RoIAwarePool3d") and r.points_in_boxes_gpu is R.points_in_boxes_gpu
        assert v1.boxes_aligned_iou3d_gpu is I1.boxes_aligned_iou3d_gpu and v1.nms_gpu() == "ref-v1" and v1.boxes_iou3d_gpu() == "ref-v1"
        shim.install(cpu_entry_points=True)
        assert m.boxes_bev_iou_cpu is I.boxes_bev_iou_cpu and r.points_in_boxes_cpu is R.points_in_boxes_cpu
    finally:
        _drop_pcdet_modules()


def test_cpu_entry_points_fail_loudly_without_a_usable_gpu(monkeypatch):
    """No host implementation exists (by design): without a GPU, or in a child forked after CUDA initialisation,
    the _cpu-named functions raise with an actionable message instead of computing on the CPU."""
    a = synth.kitti_boxes(3, 0).numpy()
    p = synth.points(10, synth.kitti_boxes(3, 0), seed=0).numpy()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            I.boxes_bev_iou_cpu(a, a)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            R.points_in_boxes_cpu(p, a)
    monkeypatch.setattr(torch.cuda, "_is_in_bad_fork", lambda: True)
    with pytest.raises(RuntimeError, match="forked after CUDA was initialised"):
        I.boxes_bev_iou_cpu(a, a)
    with pytest.raises(RuntimeError, match="spawn"):
        R.points_in_boxes_cpu(p, a)


def test_shim_registers_the_aligned_iou_module():
    _drop_pcdet_modules()
    shim.install()
    from pcdet.ops.iou3d.iou3d_utils import boxes_aligned_iou3d_gpu
    assert boxes_aligned_iou3d_gpu is I1.boxes_aligned_iou3d_gpu


def test_empty_inputs_do_not_need_a_gpu():
    assert I.boxes_bev_iou_cpu(np.zeros((0, 7), np.float32), np.zeros((3, 7), np.float32)).shape == (0, 3)
    out = R.points_in_boxes_cpu(torch.zeros((5, 3)), torch.zeros((0, 7)))
    assert out.shape == (0, 5) and out.dtype == torch.int32


def test_check_numpy_to_torch_semantics():
    x, flag = I.check_numpy_to_torch(np.zeros((2, 7), dtype=np.float64))
    assert flag and x.dtype == torch.float32
    y, flag = I.check_numpy_to_torch(torch.zeros(2, 7, dtype=torch.float64))
    assert not flag and y.dtype == torch.float64


def test_synthetic_generators_are_seeded_and_shaped():
    assert torch.equal(synth.kitti_boxes(10, 3), synth.kitti_boxes(10, 3))
    anchors = synth.anchors_kitti3()
    assert anchors.shape == (211200, 7)
    assert [round(float(v), 2) for v in anchors[:, 6].unique()] == [0.0, 1.57]
    b, s = synth.proposals(100, 5, 0)
    assert b.shape == (100, 7) and s.unique().numel() == 100
    smp, gt = synth.cvae_samples(7, 30, 0)
    assert smp.shape == (210, 7) and gt.shape == (7, 7)
    pts = synth.points(1000, synth.kitti_boxes(4, 0))
    assert pts.shape == (1000, 3) and pts.dtype == torch.float32


def test_variance_voting_nms_host_side(cpu_golden):
    """new_nms_gpu / nms_func run their loop on the device (csrc/vnms.cu; parity with the reference's Python against
    tests/golden/cpu_golden.npz is a GPU test).  What stays on the host: the heading wrap of new_nms_gpu
    (common_utils.limit_period, offset 0.5, period 2 pi, in float32) -- and no computation without a GPU."""
    from glenet_b200 import variance_nms as V
    h = torch.tensor([0.0, 3.0, 3.2, -3.2, 6.5, -9.0, 3.1415927, -3.1415927], dtype=torch.float32)
    want = h - torch.floor(h / (np.pi * 2) + 0.5) * (np.pi * 2)
    got = V._limit_period(h, offset=0.5, period=np.pi * 2)
    assert torch.equal(got, want) and float(got.abs().max()) <= np.pi + 1e-6
    if not torch.cuda.is_available():
        boxes, scores = torch.from_numpy(cpu_golden["vnms_boxes"]), torch.from_numpy(cpu_golden["vnms_scores"])
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            V.new_nms_gpu(boxes, scores, 0.25)
        with pytest.raises(RuntimeError):
            V.softnms_gpu(boxes, scores, 0.25)


def test_scale_by_iou_modes():
    from glenet_b200 import variance_nms as V
    iou = torch.tensor([0.0, 0.2, 0.5, 0.9])
    lin = V.scale_by_iou(iou, 0.3, "linear")
    assert torch.allclose(lin, torch.tensor([1.0, 1.0, 0.5, 0.1]))
    assert torch.allclose(V.scale_by_iou(iou, 0.3, "gaussian"), torch.exp(-iou ** 2 / 0.3))


def test_rotate_iou_and_crop_wrappers_host_side():
    """Argument handling that needs no device: empty problems return without touching CUDA (rotate_iou.py:303-305)."""
    import numpy as np
    from glenet_b200 import gt_database as G, rotate_iou as RI
    b = np.zeros((0, 5), dtype=np.float64)
    q = np.ones((4, 5), dtype=np.float32)
    out = RI.rotate_iou_gpu_eval(b, q)
    assert out.shape == (0, 4) and out.dtype == np.float32        # the reference returns the float32 zeros before the dtype cast
    assert RI.rotate_iou_gpu_eval(q, b).shape == (4, 0)
    assert RI.rotate_iou_gpu_eval_blocks(b, b, [0, 0], [0, 0]) [0].shape == (0, 0)
    o, c = G.crop_gt_objects(np.zeros((10, 4), dtype=np.float32), np.zeros((0, 7), dtype=np.float32), "kitti")
    assert o.tolist() == [0] and c.shape == (0, 4)
    o, c = G.crop_gt_objects(np.zeros((0, 5), dtype=np.float32), np.zeros((3, 7)), "waymo")
    assert o.tolist() == [0, 0, 0, 0] and c.shape == (0, 5)


def _spatial_groups_model(boxes):
    """numpy model of nms_spatial_kernel's grouping (csrc/nms.cu): counting sort by the Morton code of the centre's cell
    (32 x 32 cells over the finite centres), groups of <= 64 consecutive boxes confined to a 4 x 4 block of cells, the
    bounding box of each group's cull circles (infinite with a non-finite member), components of the "boxes meet" graph."""
    import numpy as np
    b = boxes.astype(np.float32)
    n = len(b)
    cx, cy = b[:, 0], b[:, 1]
    rad = (np.float32(0.5) * np.sqrt(b[:, 3] * b[:, 3] + b[:, 4] * b[:, 4])) * np.float32(1.0001) + np.float32(0.03) \
        + np.float32(2e-6) * (np.abs(cx) + np.abs(cy)) + np.float32(1e-3)
    with np.errstate(invalid="ignore", over="ignore"):
        fin = np.isfinite(cx) & np.isfinite(cy) & np.isfinite(rad)
    G = 32
    cell = np.full(n, G * G - 1, dtype=np.int64)
    if fin.any():
        x0, x1, y0, y1 = cx[fin].min(), cx[fin].max(), cy[fin].min(), cy[fin].max()
        invx = np.float32(G) / (x1 - x0) if x1 > x0 else np.float32(0)
        invy = np.float32(G) / (y1 - y0) if y1 > y0 else np.float32(0)
        ix = np.clip(((cx[fin] - x0) * invx).astype(np.int64), 0, G - 1)
        iy = np.clip(((cy[fin] - y0) * invy).astype(np.int64), 0, G - 1)
        m = np.zeros_like(ix)
        for bit in range(5):
            m |= ((ix >> bit) & 1) << (2 * bit) | ((iy >> bit) & 1) << (2 * bit + 1)
        cell[fin] = m
    perm = np.argsort(cell, kind="stable")          # (the kernel's order inside a cell is arbitrary; any order is a valid model)
    block = cell[perm] >> 4
    groups = []
    for s in range(G * G // 16):
        idx = perm[block == s]
        for k in range(0, len(idx), 64):
            groups.append(idx[k:k + 64])
    boxes_g = []
    for g in groups:
        if not fin[g].all():
            boxes_g.append((-np.inf, -np.inf, np.inf, np.inf))
        else:
            boxes_g.append(((cx[g] - rad[g]).min(), (cy[g] - rad[g]).min(), (cx[g] + rad[g]).max(), (cy[g] + rad[g]).max()))
    ng = len(groups)
    meet = np.zeros((ng, ng), dtype=bool)
    for p in range(ng):
        for q in range(ng):
            a, c = boxes_g[p], boxes_g[q]
            meet[p, q] = not (a[0] > c[2] or c[0] > a[2] or a[1] > c[3] or c[1] > a[3])
    lab = np.arange(ng)
    changed = True
    while changed:
        changed = False
        for p in range(ng):
            for q in range(ng):
                if meet[p, q] and lab[p] != lab[q]:
                    lab[p] = lab[q] = min(lab[p], lab[q]); changed = True
    return groups, meet, lab, rad - np.float32(1e-3)


def test_nms_spatial_grouping_never_separates_a_pair_the_circle_test_keeps():
    """Design invariant behind the spatial-tile NMS: the groups partition the boxes, hold at most 64 each, there are at most
    n / 64 + 64 of them, and every pair the mask kernel's circle test would keep -- including every pair with a NaN / inf
    term, which is never culled -- lies in two groups whose bounding boxes meet, hence in one component of the sweep."""
    import numpy as np
    from glenet_b200 import synth
    rng = np.random.default_rng(5)
    cases = [synth.proposals(1500, 12, 1)[0].numpy(), synth.proposals(4096, 20, 2)[0].numpy(), synth.kitti_boxes(700, 3).numpy()]
    odd = synth.proposals(900, 6, 4)[0].numpy().copy()
    odd[10, 0] = np.nan; odd[20, 3] = np.inf; odd[30:60, :2] += 4000.0; odd[100] = odd[7]; odd[200:260, :2] = odd[200, :2]
    cases.append(odd)
    cases.append(np.tile(synth.kitti_boxes(1, 9).numpy(), (300, 1)))          # every box on one spot: a single cell
    for b in cases:
        n = len(b)
        groups, meet, lab, rad = _spatial_groups_model(b)
        assert sorted(np.concatenate(groups).tolist()) == list(range(n))
        assert max(len(g) for g in groups) <= 64 and len(groups) <= (n + 63) // 64 + 64
        gid = np.empty(n, dtype=np.int64)
        for k, g in enumerate(groups):
            gid[g] = k
        sample = rng.integers(0, n, (200000, 2))
        i, j = sample[:, 0], sample[:, 1]
        cx, cy = b[:, 0].astype(np.float32), b[:, 1].astype(np.float32)
        with np.errstate(invalid="ignore", over="ignore"):
            dx, dy, rr = cx[i] - cx[j], cy[i] - cy[j], rad[i] + rad[j]
            kept = ~(dx * dx + dy * dy > rr * rr)                               # the kernel's test: NaN anywhere => not culled
        assert kept.sum() > 1000
        assert meet[gid[i[kept]], gid[j[kept]]].all()
        assert (lab[gid[i[kept]]] == lab[gid[j[kept]]]).all()
