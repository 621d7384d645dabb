"""Host-side logic that needs no GPU: shim registration, wrapper pre-launch checks, synthetic inputs."""
import importlib
import sys

import numpy as np
import pytest
import torch

from glenet_b200 import iou3d_nms_utils as I
from glenet_b200 import roiaware_pool3d_utils as R
from glenet_b200 import shim, synth


def test_shim_registers_reference_module_paths():
    shim.install()
    m = importlib.import_module("pcdet.ops.iou3d_nms.iou3d_nms_utils")
    assert m is I
    from pcdet.ops.roiaware_pool3d import roiaware_pool3d_utils as r2
    assert r2 is R
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
