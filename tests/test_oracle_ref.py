"""oracle/ref.py restates what the reference's Python wrappers do around the native calls.  Here those restatements
are checked against the reference's wrapper FILES imported verbatim (tests/golden/make_golden.py's loader), which is
only possible where /root/reference exists -- i.e. in the build container, not on the GPU box.  The GPU-dialect
wrappers cannot run here (no device); for them the source text of the reference functions is checked to be the
sequence of native calls that oracle/ref.py issues, and their results are pinned on the B200 by gpu_golden.npz."""
import importlib.util
import inspect
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT

REF = os.environ.get("GLENET_REFERENCE", "/root/reference")


@pytest.fixture(scope="module")
def ref_wrappers(ref_so):
    if not os.path.isdir(os.path.join(REF, "pcdet")):
        pytest.skip("reference tree not present (GPU box)")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN_DIR, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    saved = {k: v for k, v in sys.modules.items() if k == "pcdet" or k.startswith("pcdet.") or k == "SharedArray"}
    iou, roi = mg.load_reference_wrappers()
    yield iou, roi, mg
    for k in [k for k in sys.modules if k == "pcdet" or k.startswith("pcdet.") or k == "SharedArray"]:
        del sys.modules[k]
    sys.modules.update(saved)


def test_restated_cpu_wrappers_equal_the_reference_files(ref_wrappers, ref_so):
    iou, roi, mg = ref_wrappers
    d = mg.inputs()
    for a, b in (("sparse_a", "sparse_b"), ("dense", "dense"), ("adv", "adv"), ("waymo_p", "waymo_gt")):
        want = iou.boxes_bev_iou_cpu(d[a], d[b])
        got = ref_so.boxes_bev_iou_cpu(d[a], d[b])
        assert torch.equal(want, got), (a, b)
        # numpy in -> numpy out; the flag follows boxes_b (iou3d_nms_utils.py:61-62,68)
        w2, g2 = iou.boxes_bev_iou_cpu(d[a], d[b].numpy()), ref_so.boxes_bev_iou_cpu(d[a], d[b].numpy())
        assert isinstance(w2, np.ndarray) and isinstance(g2, np.ndarray) and np.array_equal(w2, g2)
    for f in range(2):
        pts, boxes = d["pib_points"][f], d["pib_boxes"][f]
        want = roi.points_in_boxes_cpu(pts, boxes)
        got = ref_so.points_in_boxes_cpu(pts, boxes)
        assert want.dtype == got.dtype == torch.int32 and torch.equal(want, got)
        w2, g2 = roi.points_in_boxes_cpu(pts.numpy(), boxes.numpy()), ref_so.points_in_boxes_cpu(pts.numpy(), boxes.numpy())
        assert isinstance(w2, np.ndarray) and np.array_equal(w2, g2)
        # float64 inputs are cast with .float() (roiaware_pool3d_utils.py:19-23)
        assert torch.equal(roi.points_in_boxes_cpu(pts.double(), boxes.double()), ref_so.points_in_boxes_cpu(pts.double(), boxes.double()))


def test_reference_gpu_wrappers_issue_the_calls_ref_py_restates(ref_wrappers):
    """No GPU here: pin the *text* of the reference's GPU wrappers to the call sequence oracle/ref.py issues."""
    iou, roi, _ = ref_wrappers
    src = inspect.getsource(iou.boxes_iou_bev)
    assert "boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)" in src
    src = inspect.getsource(iou.boxes_iou3d_gpu)
    for frag in ("boxes_overlap_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), overlaps_bev)", "torch.clamp(min_of_max - max_of_min, min=0)",
                 "overlaps_3d = overlaps_bev * overlaps_h", "torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)"):
        assert frag in src, frag
    src = inspect.getsource(iou.nms_gpu)
    for frag in ("scores.sort(0, descending=True)[1]", "order[:pre_maxsize]", "boxes[order].contiguous()", "iou3d_nms_cuda.nms_gpu(boxes, keep, thresh)",
                 "order[keep[:num_out].cuda()].contiguous(), None"):
        assert frag in src, frag
    src = inspect.getsource(iou.nms_normal_gpu)
    assert "iou3d_nms_cuda.nms_normal_gpu(boxes, keep, thresh)" in src
    src = inspect.getsource(roi.points_in_boxes_gpu)
    for frag in ("fill_(-1)", "roiaware_pool3d_cuda.points_in_boxes_gpu(boxes.contiguous(), points.contiguous(), box_idxs_of_pts)"):
        assert frag in src, frag
