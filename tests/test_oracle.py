"""Pin the C oracle (oracle/geom_oracle.c): against the golden vectors produced by running the
reference (tests/golden), and -- where oracle/_ref exists -- bit-for-bit against the reference's
own compiled CPU functions on fresh seeded inputs.  No GPU needed."""
import numpy as np
import pytest
import torch

from glenet_b200 import synth


def unpack(mask_packed, m):
    return np.unpackbits(mask_packed, axis=1)[:, :m].astype(np.int32)


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name,a,b", [("sparse", "sparse_a", "sparse_b"), ("dense", "dense", "dense"),
                                      ("adv", "adv", "adv"), ("waymo", "waymo_p", "waymo_gt")])
def test_iou_cpu_dialect_matches_reference_golden(cpu_golden, capi, name, a, b):
    got = capi.boxes_iou_bev(cpu_golden[a], cpu_golden[b], dialect=capi.CPU)
    want = cpu_golden[f"cpu_iou_{name}"]
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)   # bit-exact: same libm, same rounding sequence


def test_pib_cpu_dialect_matches_reference_golden(cpu_golden, capi):
    for f in range(2):
        pts, boxes = cpu_golden["pib_points"][f], cpu_golden["pib_boxes"][f]
        want = unpack(cpu_golden[f"cpu_pib_mask_{f}"], pts.shape[0])
        got = capi.points_in_boxes_mask(pts, boxes, dialect=capi.CPU)
        np.testing.assert_array_equal(got, want)
        assert want.sum() > 100   # the fixture really has points inside boxes


def test_gpu_dialect_restatement_close_to_gpu_golden(gpu_golden, cpu_golden, capi):
    """The fma-pattern restatement differs from the real GPU only by libdevice-vs-glibc trig ulps."""
    for name, a, b in (("sparse", "sparse_a", "sparse_b"), ("dense", "dense", "dense"), ("waymo", "waymo_p", "waymo_gt")):
        got = capi.boxes_iou_bev(cpu_golden[a], cpu_golden[b], dialect=capi.GPU)
        want = gpu_golden[f"gpu_iou_bev_{name}"]
        assert np.abs(got - want).max() <= 1e-5
        np.testing.assert_array_equal(got == 0, want == 0)
        got3 = capi.boxes_iou3d(cpu_golden[a], cpu_golden[b], dialect=capi.GPU)
        assert np.abs(got3 - gpu_golden[f"gpu_iou3d_{name}"]).max() <= 1e-5
    idx = capi.points_in_boxes_index(cpu_golden["pib_points"], cpu_golden["pib_boxes"], dialect=capi.GPU)
    assert (idx != gpu_golden["gpu_pib_index"]).mean() < 1e-4
    for thr in (0.7, 0.1, 0.01):
        order = np.argsort(-cpu_golden["nms_scores"], kind="stable")
        keep, near = capi.nms(cpu_golden["nms_boxes"][order], thr, normal=False, dialect=capi.GPU)
        if near == 0:
            np.testing.assert_array_equal(order[keep], gpu_golden[f"gpu_nms_{thr}"])
        keep, near = capi.nms(cpu_golden["nms_boxes"][order], thr, normal=True, dialect=capi.GPU)
        if near == 0:
            np.testing.assert_array_equal(order[keep], gpu_golden[f"gpu_nms_normal_{thr}"])


# ------------------------------------------------------------------ against the compiled reference
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_bitexact_vs_reference_cpu_iou(ref_so, capi, seed):
    a, b = synth.kitti_boxes(200, seed), synth.kitti_boxes(50, seed + 10)     # BASELINE config 0
    np.testing.assert_array_equal(capi.boxes_iou_bev(a, b), ref_so.boxes_bev_iou_cpu(a, b).numpy())
    p, _ = synth.proposals(300, 12, seed)
    want = ref_so.boxes_bev_iou_cpu(p, p).numpy()
    np.testing.assert_array_equal(capi.boxes_iou_bev(p, p), want)
    assert (want > 0).mean() > 0.03


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_bitexact_vs_reference_cpu_pib(ref_so, capi, seed):
    boxes = synth.kitti_boxes(20, seed)
    pts = synth.points(30000, boxes, synth.KITTI_RANGE, 0.2, seed=seed)
    want = ref_so.points_in_boxes_cpu(pts, boxes).numpy()
    np.testing.assert_array_equal(capi.points_in_boxes_mask(pts, boxes), want)
    assert want.sum() > 1000


def test_reference_wrapper_quirks(ref_so):
    """numpy-in/numpy-out follows the LAST converted argument (iou3d_nms_utils.py:61-62,68)."""
    a, b = synth.kitti_boxes(4, 0), synth.kitti_boxes(3, 1)
    assert isinstance(ref_so.boxes_bev_iou_cpu(a.numpy(), b.numpy()), np.ndarray)
    assert isinstance(ref_so.boxes_bev_iou_cpu(a.numpy(), b), torch.Tensor)
    assert isinstance(ref_so.boxes_bev_iou_cpu(a, b.numpy()), np.ndarray)


# ------------------------------------------------------------------ oracle self-consistency
def test_oracle_properties(capi):
    a = synth.kitti_boxes(64, 3)
    iou = capi.boxes_iou_bev(a, a)
    assert np.allclose(np.diag(iou), 1.0, atol=1e-5)
    assert np.abs(iou - iou.T).max() < 1e-5
    # first-hit index == argmax of the mask's first set row
    boxes = synth.kitti_boxes(12, 5)
    pts = synth.points(5000, boxes, synth.KITTI_RANGE, 0.4, seed=5)
    mask = capi.points_in_boxes_mask(pts, boxes, dialect=capi.GPU)
    idx = capi.points_in_boxes_index(pts[None], boxes[None], dialect=capi.GPU)[0]
    first = np.where(mask.any(0), mask.argmax(0), -1)
    np.testing.assert_array_equal(idx, first)
    # NMS keeps are sorted, start with 0 and are idempotent
    p, s = synth.proposals(300, 10, 1)
    order = np.argsort(-s.numpy(), kind="stable")
    keep, _ = capi.nms(p.numpy()[order], 0.5)
    assert keep[0] == 0 and np.all(np.diff(keep) > 0)
    keep2, _ = capi.nms(p.numpy()[order][keep], 0.5)
    np.testing.assert_array_equal(keep2, np.arange(len(keep)))


def test_flop_model_counts(capi):
    p, _ = synth.proposals(100, 5, 0)
    _, st = capi.boxes_iou_bev(p, p, dialect=capi.GPU, stats=True)
    assert st.pairs == 100 * 100
    assert sum(st.cnt_hist) == st.pairs
    assert st.flops() > 151 * st.pairs


# ------------------------------------------------------------------ pcdet/ops/iou3d (boxes_aligned_iou3d_gpu's op)
def test_v1_overlap_cpu_dialect_matches_reference_golden(cpu_golden_v1, capi):
    g = cpu_golden_v1
    np.testing.assert_array_equal(capi.iou3d_v1_overlap_aligned(g["v1_pred_bev"], g["v1_tgt_bev"], dialect=capi.CPU), g["cpu_v1_overlap_aligned"])
    np.testing.assert_array_equal(capi.iou3d_v1_overlap_bev(g["v1_pred_bev"][:64], g["v1_tgt_bev"][:64], dialect=capi.CPU), g["cpu_v1_overlap_block"])
    assert (g["cpu_v1_overlap_aligned"] > 0).mean() > 0.8 and (g["cpu_v1_overlap_aligned"] == 0).sum() > 20   # overlapping pairs and misses


@pytest.mark.parametrize("seed", [1, 2])
def test_v1_oracle_bitexact_vs_reference_cpu(ref_iou3d, capi, seed):
    pred, tgt = synth.head_pairs(300, seed)
    a5, b5 = ref_iou3d.boxes3d_to_bev_torch(pred), ref_iou3d.boxes3d_to_bev_torch(tgt)
    want = ref_iou3d.iou3d_v1_overlap_bev_cpu(a5, b5[:120]).numpy()
    np.testing.assert_array_equal(capi.iou3d_v1_overlap_bev(a5.numpy(), b5[:120].numpy(), dialect=capi.CPU), want)
    np.testing.assert_array_equal(capi.iou3d_v1_overlap_aligned(a5.numpy()[:120], b5.numpy()[:120], dialect=capi.CPU), want.diagonal())


def test_v1_restated_python_wrapper_equals_reference_file(ref_iou3d):
    """oracle.ref.boxes3d_to_bev_torch restates pcdet/ops/iou3d/iou3d_utils.py:79-106; check it against the file itself."""
    import os
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree not present")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    u = make_golden.load_reference_iou3d_utils()
    pred, _ = synth.head_pairs(50, 5)
    for mode in ("wlh", "lwh", "hwl"):
        assert torch.equal(u.boxes3d_to_bev_torch(pred, mode), ref_iou3d.boxes3d_to_bev_torch(pred, mode))


# ------------------------------------------------------------------ next scope row: KITTI evaluator's rotated IoU
def test_rotate_iou_eval_oracle_vs_reference_golden():
    """oracle/rotate_iou_oracle.c against the reference's own numba kernel, run by numba's CUDA simulator
    (tests/golden/make_golden_rotate_iou.py).  Tolerance 1e-5 absolute: the simulator and the restatement differ in
    float32 / float64 promotion of a few intermediates; zeros must be the same zeros."""
    import os
    from conftest import ROOT
    from oracle import rotate_iou
    g = np.load(os.path.join(ROOT, "tests", "golden", "rotate_iou_golden.npz"))
    for crit in (-1, 0, 1):
        got, want = rotate_iou.rotate_iou_eval(g["boxes"], g["query"], crit), g[f"iou_{crit}"]
        assert got.shape == want.shape == (70, 45) and got.dtype == np.float32
        assert np.abs(got - want).max() <= 1e-5
        np.testing.assert_array_equal(got == 0, want == 0)
    assert (g["iou_-1"] > 0).sum() > 400
    # criterion semantics (rotate_iou.py:249-261): the QUERY box is rbox1 of devRotateIoUEval (:281-283)
    inter = rotate_iou.rotate_iou_eval(g["boxes"], g["query"], 2)
    a_q, a_b = g["query"][:, 2] * g["query"][:, 3], g["boxes"][:, 2] * g["boxes"][:, 3]
    np.testing.assert_allclose(rotate_iou.rotate_iou_eval(g["boxes"], g["query"], 0), inter / a_q[None, :], rtol=0, atol=1e-6)
    np.testing.assert_allclose(rotate_iou.rotate_iou_eval(g["boxes"], g["query"], 1), inter / a_b[:, None], rtol=0, atol=1e-6)
    assert rotate_iou.rotate_iou_eval(g["boxes"][:0], g["query"]).shape == (0, 45)


def test_rotate_iou_eval_oracle_contraction_dialect():
    """Dialect 1 (the FMA pattern of the kernel numba compiles for sm_100a) stays within 1e-5 of the simulator goldens, and
    -- with the cos / sin tables libdevice produced -- reproduces the goldens the reference kernel wrote on a B200 bit for bit."""
    import os
    from conftest import ROOT
    from oracle import rotate_iou
    g = np.load(os.path.join(ROOT, "tests", "golden", "rotate_iou_golden.npz"))
    for crit in (-1, 0, 1):
        got = rotate_iou.rotate_iou_eval(g["boxes"], g["query"], crit, contract=True)
        assert np.abs(got - g[f"iou_{crit}"]).max() <= 1e-5
    path = os.path.join(ROOT, "tests", "golden", "rotate_iou_gpu_golden.npz")
    if not os.path.isfile(path):
        pytest.skip("tests/golden/rotate_iou_gpu_golden.npz missing (make_golden_rotate_iou.py gpu on a B200)")
    h = np.load(path)
    for crit in (-1, 0, 1, 2):
        got = rotate_iou.rotate_iou_eval(h["boxes"], h["query"], crit, contract=True, trig_boxes=h["trig_boxes"], trig_query=h["trig_query"])
        want = h[f"iou_{crit}"]
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        assert same.all(), (crit, int((~same).sum()), float(np.nanmax(np.abs(got - want))))
