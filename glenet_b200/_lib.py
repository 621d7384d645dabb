"""ctypes binding of ``libglenet_geom.so`` (the C ABI declared in ``include/glenet_geom.h``).

There is no CPU or eager fallback: if the shared object is missing it is built with nvcc
(``glenet_b200.build``); if that is impossible the import of the op fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None

c_float_p = ctypes.c_void_p  # raw device pointers travel as integers (tensor.data_ptr())

_SIGNATURES = {
    "glenet_abi_version": (ctypes.c_int, []),
    "glenet_last_error": (ctypes.c_char_p, []),
    "glenet_boxes_overlap_bev_gpu": (ctypes.c_int, [c_float_p, ctypes.c_int, c_float_p, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "glenet_boxes_iou_bev_gpu": (ctypes.c_int, [c_float_p, ctypes.c_int, c_float_p, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "glenet_boxes_iou3d_gpu": (ctypes.c_int, [c_float_p, ctypes.c_int, c_float_p, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "glenet_boxes_iou_frames_gpu": (ctypes.c_int, [ctypes.c_int, c_float_p, ctypes.c_longlong, ctypes.c_int, c_float_p, ctypes.c_longlong,
                                                   ctypes.c_int, c_float_p, ctypes.c_int, ctypes.c_void_p]),
    "glenet_boxes_iou_frames_sparse_gpu": (ctypes.c_int, [ctypes.c_int, c_float_p, ctypes.c_longlong, ctypes.c_int, c_float_p, ctypes.c_longlong,
                                                          ctypes.c_int, ctypes.c_int, ctypes.c_void_p, c_float_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p]),
    "glenet_boxes_iou_frames_max_gpu": (ctypes.c_int, [ctypes.c_int, c_float_p, ctypes.c_longlong, ctypes.c_int, c_float_p, ctypes.c_longlong,
                                                       ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "glenet_iou_keys_decode_gpu": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong,
                                                  c_float_p, ctypes.c_void_p, c_float_p, ctypes.c_void_p, ctypes.c_void_p]),
    "glenet_symm_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "glenet_symm_free": (ctypes.c_int, [ctypes.c_void_p]),
    "glenet_symm_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    "glenet_symm_import": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "glenet_symm_unmap": (ctypes.c_int, [ctypes.c_void_p]),
    "glenet_exchange_window_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_longlong]),
    "glenet_exchange_status": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint)]),
    "glenet_boxes_iou_frames_assign_gpu": (ctypes.c_int, [ctypes.c_int, c_float_p, ctypes.c_longlong, ctypes.c_int, c_float_p, ctypes.c_longlong,
                                                          ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p,
                                                          c_float_p, ctypes.c_void_p, c_float_p, ctypes.c_void_p,
                                                          ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_longlong, ctypes.c_uint, ctypes.c_void_p]),
    "glenet_boxes_iou_frames_gather_gpu": (ctypes.c_int, [ctypes.c_int, c_float_p, ctypes.c_longlong, ctypes.c_int, c_float_p, ctypes.c_longlong,
                                                          ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_longlong,
                                                          ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_longlong, ctypes.c_uint, ctypes.c_void_p]),
    "glenet_iou3d_v1_boxes_aligned_gpu": (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                         c_float_p, c_float_p, c_float_p, ctypes.c_void_p]),
    "glenet_iou3d_v1_aligned_overlap_bev_gpu": (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "glenet_iou3d_v1_aligned_overlap_bev_cpu_dialect": (ctypes.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "glenet_boxes_iou_aligned_gpu": (ctypes.c_int, [ctypes.c_int, c_float_p, ctypes.c_int, c_float_p, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "glenet_boxes_iou_bev_cpu_dialect": (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, c_float_p, c_float_p, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "glenet_host_trig4": (None, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "glenet_host_trig4_strided": (None, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "glenet_host_trig2": (None, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "glenet_nms_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "glenet_nms_gpu": (ctypes.c_int, [c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "glenet_nms_normal_gpu": (ctypes.c_int, [c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "glenet_variance_nms_gpu": (ctypes.c_int, [c_float_p, c_float_p, c_float_p, ctypes.c_int, c_float_p, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]),
    "glenet_points_in_boxes_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "glenet_points_in_boxes_gpu": (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "glenet_points_in_boxes_cpu_dialect": (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, c_float_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "glenet_gt_crop_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_longlong]),
    "glenet_gt_crop_gpu": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, c_float_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_longlong, ctypes.c_void_p, c_float_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "glenet_rotate_iou_eval_gpu": (ctypes.c_int, [c_float_p, ctypes.c_int, c_float_p, ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "glenet_cvae_iou3d_gpu": (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "glenet_rotate_iou_eval_blocks_gpu": (ctypes.c_int, [c_float_p, ctypes.c_void_p, c_float_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)
ABI_VERSION = 13


def lib_path() -> str:
    return _build.LIBPATH


def load() -> ctypes.CDLL:
    """Load (building first if needed) the native library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIBPATH
        # (re)build when the .so is missing or older than any source; returns at once when nothing is stale.  On a box
        # without nvcc an existing .so is used as it is.
        try:
            _build.build_library()
        except RuntimeError:
            if not os.path.isfile(path):
                raise
        lib = ctypes.CDLL(path)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        got = lib.glenet_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"libglenet_geom.so ABI {got} != expected {ABI_VERSION}: rebuild with `python -m glenet_b200.build --force`")
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().glenet_last_error().decode(errors="replace")
        raise RuntimeError(f"libglenet_geom {what} failed ({rc}): {msg}")
