"""Register glenet_b200 under the reference's module paths.

    import glenet_b200.shim; glenet_b200.shim.install()
    from pcdet.ops.iou3d_nms import iou3d_nms_utils          # -> glenet_b200.iou3d_nms_utils
    from pcdet.ops.roiaware_pool3d import roiaware_pool3d_utils
    from pcdet.ops.iou3d.iou3d_utils import boxes_aligned_iou3d_gpu   # -> glenet_b200.iou3d_utils

Inside a real OpenPCDet/GLENet checkout the two modules are replaced in ``sys.modules`` (call
``install()`` before anything imports ``pcdet.ops``); without pcdet installed, stub parent
packages are created so that the reference's import statements work unchanged
(``pcdet/models/model_utils/model_nms_utils.py:3``, ``pcdet/utils/box_utils.py:6``).
"""
from __future__ import annotations

import importlib
import sys
import types

from . import iou3d_nms_utils, iou3d_utils, roiaware_pool3d_utils

_TARGETS = {
    "pcdet.ops.iou3d_nms.iou3d_nms_utils": iou3d_nms_utils,
    "pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils": roiaware_pool3d_utils,
    # only boxes_aligned_iou3d_gpu (+ its two helpers) of pcdet/ops/iou3d is provided: the one function GLENet imports from it
    "pcdet.ops.iou3d.iou3d_utils": iou3d_utils,
}


def _ensure_package(name: str):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        mod = types.ModuleType(name)
        mod.__path__ = []  # mark as package
        sys.modules[name] = mod
        if "." in name:
            parent, _, leaf = name.rpartition(".")
            setattr(_ensure_package(parent), leaf, mod)
        return mod


def install() -> None:
    for dotted, mod in _TARGETS.items():
        parent, _, leaf = dotted.rpartition(".")
        pkg = _ensure_package(parent)
        sys.modules[dotted] = mod
        setattr(pkg, leaf, mod)
