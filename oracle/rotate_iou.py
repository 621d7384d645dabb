"""ctypes binding of ``oracle/librotate_iou_oracle.so`` -- TEST INFRASTRUCTURE ONLY.

CPU restatement of ``rotate_iou_gpu_eval`` (pcdet/datasets/kitti/kitti_object_eval_python/rotate_iou.py:263-330), the
rotated IoU of the KITTI evaluator; checks ``glenet_b200.rotate_iou`` (csrc/rotate_iou.cu)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "librotate_iou_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "rotate_iou_oracle.c")
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "librotate_iou_oracle.so"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB


def rotate_iou_eval(boxes: np.ndarray, query_boxes: np.ndarray, criterion: int = -1, contract: bool = False,
                    trig_boxes: np.ndarray = None, trig_query: np.ndarray = None) -> np.ndarray:
    """(N, 5) x (K, 5) [x, y, x_d, y_d, angle] -> (N, K) float32; criterion -1: IoU, 0: / area(query), 1: / area(box).

    ``contract``: the FMA contraction of the numba kernel on sm_100a (see the C file).  ``trig_*``: optional (n, 2) tables of
    {cos, sin}(angle) as libdevice computed them -- host libm otherwise (differs from libdevice in the last bit for some angles)."""
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_rotate_iou_eval_dialect.restype = None
        _lib.oracle_rotate_iou_eval_dialect.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    b = np.ascontiguousarray(boxes, dtype=np.float32)
    q = np.ascontiguousarray(query_boxes, dtype=np.float32)
    assert b.ndim == 2 and b.shape[1] == 5 and q.ndim == 2 and q.shape[1] == 5
    out = np.zeros((b.shape[0], q.shape[0]), dtype=np.float32)
    tb = None if trig_boxes is None else np.ascontiguousarray(trig_boxes, dtype=np.float32)
    tq = None if trig_query is None else np.ascontiguousarray(trig_query, dtype=np.float32)
    assert tb is None or tb.shape == (b.shape[0], 2)
    assert tq is None or tq.shape == (q.shape[0], 2)
    if out.size:
        _lib.oracle_rotate_iou_eval_dialect(b.ctypes.data, b.shape[0], q.ctypes.data, q.shape[0], int(criterion), out.ctypes.data, int(bool(contract)),
                                            None if tb is None else tb.ctypes.data, None if tq is None else tq.ctypes.data)
    return out
