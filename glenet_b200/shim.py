"""Register glenet_b200 under the reference's module paths.

    import glenet_b200.shim; glenet_b200.shim.install()
    from pcdet.ops.iou3d_nms import iou3d_nms_utils          # -> functions of glenet_b200.iou3d_nms_utils
    from pcdet.ops.roiaware_pool3d import roiaware_pool3d_utils
    from pcdet.ops.iou3d.iou3d_utils import boxes_aligned_iou3d_gpu   # -> glenet_b200.iou3d_utils

Two situations:

* **Inside a real OpenPCDet / GLENet checkout** (the reference module imports): the original module stays in
  ``sys.modules`` and only the functions this package provides are patched onto it with ``setattr`` -- everything
  else the module exports (``RoIAwarePool3d`` used by ``pcdet/models/roi_heads/partA2_head.py``, the other functions
  of ``pcdet.ops.iou3d.iou3d_utils``) keeps working.  The two ``_cpu`` entry points (``boxes_bev_iou_cpu``,
  ``points_in_boxes_cpu``) are NOT patched by default: the reference calls them from forked DataLoader workers
  (``pcdet/datasets/augmentor/database_sampler.py:246-247``, ``pcdet/utils/box_utils.py:86``,
  ``pcdet/datasets/augmentor/augmentor_utils.py:149``), where CUDA cannot be initialised once the parent has touched
  it; this package has no host implementation (by design: no CPU fallback), so those call sites keep the
  reference's own CPU code unless ``install(cpu_entry_points=True)`` is requested (main-process or spawn-worker use:
  GT-database creation, ``new_nms_gpu``).
* **Without pcdet installed**: stub parent packages are created so that the reference's import statements work
  unchanged (``pcdet/models/model_utils/model_nms_utils.py:3``, ``pcdet/utils/box_utils.py:6``); all provided
  functions are present, the ``_cpu`` ones included.
"""
from __future__ import annotations

import importlib
import sys
import types

from . import cvae_eval_utils, iou3d_nms_utils, iou3d_utils, roiaware_pool3d_utils, rotate_iou

_TARGETS = {
    "pcdet.ops.iou3d_nms.iou3d_nms_utils": iou3d_nms_utils,
    "pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils": roiaware_pool3d_utils,
    # only boxes_aligned_iou3d_gpu (+ its helpers) of pcdet/ops/iou3d is provided: the one function GLENet imports from it
    "pcdet.ops.iou3d.iou3d_utils": iou3d_utils,
    # the KITTI evaluator's rotated IoU (a numba-CUDA module in the reference) ...
    "pcdet.datasets.kitti.kitti_object_eval_python.rotate_iou": rotate_iou,
    # ... and the recall IoU of the CVAE evaluation (cvae_uncertainty/ is run from its own directory: `eval_utils.eval_utils`)
    "eval_utils.eval_utils": cvae_eval_utils,
}
# modules that bind one of the patched functions with `from x import f` at import time (eval.py:5: from .rotate_iou import
# rotate_iou_gpu_eval): the name is rebound there as well when the module is already loaded
_REBIND = {"pcdet.datasets.kitti.kitti_object_eval_python.eval": ("rotate_iou_gpu_eval", "bev_box_overlap", "d3_box_overlap")}
# host-signature functions that execute on the GPU here (see the module docstring)
CPU_ENTRY_POINTS = ("boxes_bev_iou_cpu", "points_in_boxes_cpu", "boxes_aligned_overlap_bev_cpu")


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


def _import_original(dotted: str):
    """The reference's own module if it is importable (a real pcdet checkout with built extensions), else None."""
    mod = sys.modules.get(dotted)
    if mod is not None:
        return None if getattr(mod, "__glenet_b200_stub__", False) or mod in _TARGETS.values() else mod
    try:
        return importlib.import_module(dotted)
    except Exception:
        return None


def install(cpu_entry_points: bool | None = None) -> dict:
    """Patch / register the drop-in functions.  Returns ``{module path: [names patched]}``.

    ``cpu_entry_points``: patch the ``_cpu``-named functions too.  Default: only when the reference module is absent
    (stub mode); inside a real checkout they keep the reference's host code (fork-safe DataLoader workers)."""
    patched = {}
    for dotted, ours in _TARGETS.items():
        names = list(getattr(ours, "__all__"))
        orig = _import_original(dotted)
        if orig is not None:
            with_cpu = bool(cpu_entry_points)
            done = []
            for name in names:
                if name in CPU_ENTRY_POINTS and not with_cpu:
                    continue
                setattr(orig, name, getattr(ours, name))
                done.append(name)
            patched[dotted] = done
            for user, fnames in _REBIND.items():
                mod = sys.modules.get(user)
                for fname in fnames:
                    if mod is not None and fname in names and hasattr(mod, fname):
                        setattr(mod, fname, getattr(ours, fname))
            continue
        # stub mode: a fresh module that carries exactly the provided functions
        parent, _, leaf = dotted.rpartition(".")
        pkg = _ensure_package(parent)
        mod = types.ModuleType(dotted)
        mod.__glenet_b200_stub__ = True
        mod.__doc__ = ours.__doc__
        with_cpu = True if cpu_entry_points is None else bool(cpu_entry_points)
        done = []
        for name in names:
            if name in CPU_ENTRY_POINTS and not with_cpu:
                continue
            setattr(mod, name, getattr(ours, name))
            done.append(name)
        mod.__all__ = done
        sys.modules[dotted] = mod
        setattr(pkg, leaf, mod)
        patched[dotted] = done
    return patched
