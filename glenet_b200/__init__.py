"""glenet_b200: B200-native (sm_100a) rotated-box geometry hot path of GLENet / OpenPCDet.

    from glenet_b200 import iou3d_nms_utils, roiaware_pool3d_utils

are drop-ins for ``pcdet.ops.iou3d_nms.iou3d_nms_utils`` and the points-in-boxes half of
``pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils``; ``glenet_b200.shim.install()`` registers
them under the reference's module paths.  All compute goes through ``libglenet_geom.so``
(C ABI in ``include/glenet_geom.h``); nothing here falls back to the CPU.
"""
from . import iou3d_nms_utils, iou3d_utils, roiaware_pool3d_utils  # noqa: F401
from ._lib import EXPORTS, lib_path, load  # noqa: F401

__version__ = "0.1.0"
