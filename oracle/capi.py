"""ctypes binding of ``oracle/libgeom_oracle.so`` (plain-C restatement) -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libgeom_oracle.so")
CPU, GPU = 0, 1   # dialects


class IouStats(ctypes.Structure):
    _fields_ = [("pairs", ctypes.c_uint64), ("k1", ctypes.c_uint64), ("k2", ctypes.c_uint64),
                ("corners", ctypes.c_uint64), ("cnt_hist", ctypes.c_uint64 * 25)]

    def flops(self) -> float:
        """Algorithmic FLOPs of the evaluated pairs (F_pair heavy-path formula)."""
        f = self.pairs * (64 + 80 + 7) + 32 * self.k1 + 19 * self.k2 + 2 * self.corners
        for cnt, num in enumerate(self.cnt_hist):
            if cnt > 0 and num:
                f += num * (27 * cnt + cnt * (cnt - 1) // 2 + 8 * (cnt - 1))
        return float(f)


_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "geom_oracle.c")
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libgeom_oracle.so"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB)
        fp, ip, i64p = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
        lib.oracle_boxes_iou_bev.argtypes = [ctypes.c_int, fp, ctypes.c_int, fp, ctypes.c_int, fp, ctypes.POINTER(IouStats)]
        lib.oracle_boxes_iou_bev_flagged.argtypes = [ctypes.c_int, fp, ctypes.c_int, fp, ctypes.c_int, fp, ctypes.c_void_p]
        lib.oracle_boxes_iou_bev_rows.argtypes = [ctypes.c_int, fp, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int, fp]
        lib.oracle_boxes_overlap_bev.argtypes = [ctypes.c_int, fp, ctypes.c_int, fp, ctypes.c_int, fp]
        lib.oracle_boxes_iou3d.argtypes = [ctypes.c_int, fp, ctypes.c_int, fp, ctypes.c_int, fp]
        lib.oracle_nms.argtypes = [ctypes.c_int, ctypes.c_int, fp, ctypes.c_int, ctypes.c_float, i64p, ctypes.c_float, ctypes.POINTER(ctypes.c_int)]
        lib.oracle_nms.restype = ctypes.c_int
        lib.oracle_points_in_boxes_mask.argtypes = [ctypes.c_int, fp, ctypes.c_int, fp, ctypes.c_int, ip]
        lib.oracle_points_in_boxes_mask_rows.argtypes = [ctypes.c_int, fp, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int, ip]
        lib.oracle_points_in_boxes_index.argtypes = [ctypes.c_int, fp, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int, ip]
        lib.oracle_count_circle_pass.argtypes = [fp, ctypes.c_int, fp, ctypes.c_int]
        lib.oracle_iou3d_v1_overlap_bev.argtypes = [ctypes.c_int, fp, ctypes.c_int, fp, ctypes.c_int, fp]
        lib.oracle_iou3d_v1_overlap_aligned.argtypes = [ctypes.c_int, fp, fp, ctypes.c_int, fp]
        lib.oracle_count_circle_pass.restype = ctypes.c_uint64
        lib.oracle_abi_version.restype = ctypes.c_int
        _lib = lib
    return _lib


def _f32(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


def boxes_iou_bev(a, b, dialect=CPU, stats=False):
    a, b = _f32(a), _f32(b)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    st = IouStats()
    load().oracle_boxes_iou_bev(dialect, a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], out.ctypes.data, ctypes.byref(st))
    return (out, st) if stats else out


def boxes_iou_bev_flagged(a, b, dialect=GPU):
    """Returns (iou, near) where near marks pairs whose margin predicate is within 1e-4 m of flipping."""
    a, b = _f32(a), _f32(b)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    near = np.zeros((a.shape[0], b.shape[0]), dtype=np.uint8)
    load().oracle_boxes_iou_bev_flagged(dialect, a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], out.ctypes.data, near.ctypes.data)
    return out, near.astype(bool)


def boxes_overlap_bev(a, b, dialect=GPU):
    a, b = _f32(a), _f32(b)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    load().oracle_boxes_overlap_bev(dialect, a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], out.ctypes.data)
    return out


def boxes_iou3d(a, b, dialect=GPU):
    a, b = _f32(a), _f32(b)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    load().oracle_boxes_iou3d(dialect, a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], out.ctypes.data)
    return out


def nms(boxes_sorted, thresh, normal=False, dialect=GPU, tol=1e-5):
    """Returns (keep indices into boxes_sorted, number of pairs whose IoU is within tol of thresh)."""
    b = _f32(boxes_sorted)
    keep = np.zeros((b.shape[0],), dtype=np.int64)
    near = ctypes.c_int(0)
    n = load().oracle_nms(dialect, int(normal), b.ctypes.data, b.shape[0], float(thresh), keep.ctypes.data, float(tol), ctypes.byref(near))
    return keep[:n].copy(), near.value


def points_in_boxes_mask(points, boxes, dialect=CPU):
    p, b = _f32(points), _f32(boxes)
    out = np.zeros((b.shape[0], p.shape[0]), dtype=np.int32)
    load().oracle_points_in_boxes_mask(dialect, b.ctypes.data, b.shape[0], p.ctypes.data, p.shape[0], out.ctypes.data)
    return out


def points_in_boxes_index(points, boxes, dialect=GPU):
    """points (B, M, 3), boxes (B, N, 7) -> (B, M) int32 first-hit index."""
    p, b = _f32(points), _f32(boxes)
    out = np.zeros((p.shape[0], p.shape[1]), dtype=np.int32)
    load().oracle_points_in_boxes_index(dialect, b.ctypes.data, p.shape[0], b.shape[1], p.ctypes.data, p.shape[1], out.ctypes.data)
    return out


def count_circle_pass(a, b):
    a, b = _f32(a), _f32(b)
    return int(load().oracle_count_circle_pass(a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0]))


def iou3d_v1_overlap_bev(a5, b5, dialect=CPU):
    """pcdet/ops/iou3d boxes_overlap_bev_cpu: (N, 5) x (M, 5) [x1, y1, x2, y2, angle] -> (N, M)."""
    a5, b5 = _f32(a5), _f32(b5)
    out = np.zeros((a5.shape[0], b5.shape[0]), dtype=np.float32)
    load().oracle_iou3d_v1_overlap_bev(dialect, a5.ctypes.data, a5.shape[0], b5.ctypes.data, b5.shape[0], out.ctypes.data)
    return out


def iou3d_v1_overlap_aligned(a5, b5, dialect=GPU):
    """pcdet/ops/iou3d boxes_aligned_overlap_bev_gpu: row i of a5 against row i of b5 -> (N,)."""
    a5, b5 = _f32(a5), _f32(b5)
    assert a5.shape == b5.shape
    out = np.zeros((a5.shape[0],), dtype=np.float32)
    load().oracle_iou3d_v1_overlap_aligned(dialect, a5.ctypes.data, b5.ctypes.data, a5.shape[0], out.ctypes.data)
    return out
