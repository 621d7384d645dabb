"""The phased clip of csrc/clip.cuh, compiled for the HOST (tests/emul) and walked through on the CPU, against the C oracle
in the GPU dialect (oracle/geom_oracle.c, the restatement of iou3d_nms_kernel.cu:104-234 pinned to the reference).

This is the no-GPU check of the clip's logic -- slot assignment from the result bits, round-robin dealing of the crossings
to the quad lanes, packed-key sorting network, the > 8-vertex path.  Both sides get their trigonometry from glibc, so the
margin predicate and every crossing point are bit-identical and any difference comes from the clip's own structure.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT
from glenet_b200 import synth

EMUL_DIR = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libclip_emul.so")
    srcs = [os.path.join(EMUL_DIR, "clip_emul.cpp"), os.path.join(EMUL_DIR, "cuda_shim.h"),
            os.path.join(ROOT, "glenet_b200", "csrc", "clip.cuh"), os.path.join(ROOT, "glenet_b200", "csrc", "geom.cuh")]
    if not os.path.isfile(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", so, srcs[0]], check=True)
    lib = ctypes.CDLL(so)
    lib.emul_clip_aligned.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] + [ctypes.c_void_p] * 3
    lib.emul_clip_reference_chain.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p]
    lib.emul_clip_aligned_lane.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p]
    lib.emul_overlap_approx.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] + [ctypes.c_void_p] * 4
    return lib


def trig4(boxes):
    h = boxes[:, 6].astype(np.float32)
    lib = ctypes.CDLL("libm.so.6")
    lib.cosf.restype = lib.sinf.restype = ctypes.c_float
    lib.cosf.argtypes = lib.sinf.argtypes = [ctypes.c_float]
    out = np.empty((len(h), 4), dtype=np.float32)
    for i, v in enumerate(h):
        out[i] = (lib.cosf(v), lib.sinf(v), lib.cosf(-v), lib.sinf(-v))
    return out


def run(emul, a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    n = a.shape[0]
    ta, tb = trig4(a), trig4(b)
    ov, iou, cnt = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.int32)
    emul.emul_clip_aligned(a.ctypes.data, ta.ctypes.data, b.ctypes.data, tb.ctypes.data, n, ov.ctypes.data, iou.ctypes.data, cnt.ctypes.data)
    chain = np.empty(n, np.float32)
    emul.emul_clip_reference_chain(a.ctypes.data, ta.ctypes.data, b.ctypes.data, tb.ctypes.data, n, chain.ctypes.data)
    lane_variant = np.empty(n, np.float32)
    emul.emul_clip_aligned_lane(a.ctypes.data, ta.ctypes.data, b.ctypes.data, tb.ctypes.data, n, lane_variant.ctypes.data)
    np.testing.assert_array_equal(ov.view(np.uint32), lane_variant.view(np.uint32))    # quad variant == lane variant, bit for bit (NaN included)
    return ov, iou, cnt, chain


def oracle_aligned(capi, a, b):
    """Row-aligned IoU / overlap from the oracle's pairwise functions (blocks of 64 rows, diagonal taken)."""
    n = a.shape[0]
    iou, ov = np.empty(n, np.float32), np.empty(n, np.float32)
    for r0 in range(0, n, 64):
        r1 = min(n, r0 + 64)
        iou[r0:r1] = np.diagonal(capi.boxes_iou_bev(a[r0:r1], b[r0:r1], dialect=capi.GPU))
        ov[r0:r1] = np.diagonal(capi.boxes_overlap_bev(a[r0:r1], b[r0:r1], dialect=capi.GPU))
    return iou, ov


@pytest.mark.parametrize("name", ["cvae", "proposals", "anchors", "adversarial"])
def test_phased_clip_matches_the_oracle(emul, capi, cpu_golden, name):
    if name == "cvae":                                   # cfg3: samples around their GT (every pair overlaps, 1.3 % with > 8 vertices)
        smp, gt = synth.cvae_samples(400, 30, 1)
        a, b = smp.numpy(), gt.repeat_interleave(30, dim=0).numpy()
    elif name == "proposals":                            # cfg1-like clusters, all pairs of 160 proposals
        p = torch.from_numpy(cpu_golden["dense"])
        a, b = p.repeat_interleave(160, dim=0).numpy(), p.repeat(160, 1).numpy()
    elif name == "anchors":                              # axis-aligned / quarter-turn anchors near random GT
        gt = synth.kitti_boxes(100, 4)
        anc = synth.anchors_kitti3()
        d = (anc[:, None, :2] - gt[None, :, :2]).norm(dim=2)
        ia, ib = torch.nonzero(d < 3.0, as_tuple=True)
        a, b = anc[ia[:20000]].numpy(), gt[ib[:20000]].numpy()
    else:                                                # identical boxes, shared edges, corners 0.01 +- ulp from an edge, zero boxes, huge headings
        adv = torch.from_numpy(cpu_golden["adv"])
        n = adv.shape[0]
        a, b = adv.repeat_interleave(n, dim=0).numpy(), adv.repeat(n, 1).numpy()
    ov, iou, cnt, chain = run(emul, a, b)
    want_iou, want_ov = oracle_aligned(capi, a, b)
    ok = ~(np.isnan(want_ov) | np.isnan(ov))
    np.testing.assert_array_equal(np.isnan(want_ov), np.isnan(ov))
    assert np.abs(ov[ok] - want_ov[ok]).max() <= 2e-5 * max(1.0, float(np.abs(want_ov[ok]).max()))
    okk = ~np.isnan(want_iou)
    assert np.abs(iou[okk] - want_iou[okk]).max() <= 1e-5
    np.testing.assert_array_equal(ov[ok] == 0, want_ov[ok] == 0)              # exact zeros are part of the contract
    exact = (ov[ok] == want_ov[ok]).mean()
    assert exact >= (0.97 if name == "adversarial" else 0.9995), exact         # ordering ties aside, bit-identical
    # the phased structure and the round-1 single chain are the same arithmetic
    assert (ov[ok] == chain[ok]).mean() >= (0.97 if name == "adversarial" else 0.9995)
    if name == "cvae":
        assert (cnt > 8).mean() > 0.005 and (cnt >= 3).mean() > 0.99           # the slow path is exercised


def test_near_collinear_vertices_keep_their_exact_order(emul, capi):
    """Regression (found on the B200 against the reference kernel, cfg3 pairs 311770 / 344246 / ...): a corner admitted by the
    0.01 m margin and a crossing next to it can lie on one ray from the centroid; their pseudo-angle keys then differ in the
    last bits only and the packed-key network must not treat them as interchangeable (1e-3 of IoU).  ``ref_gpu_iou_bev`` are
    the values of the reference's CUDA kernel recorded on the GPU box."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "clip_regress.npz"))
    ov, iou, cnt, chain = run(emul, d["a"], d["b"])
    want_iou, want_ov = oracle_aligned(capi, d["a"], d["b"])
    assert np.abs(iou - want_iou).max() <= 1e-6
    assert np.abs(iou - d["ref_gpu_iou_bev"]).max() <= 1e-6       # libdevice vs glibc trigonometry: last-bit differences only
    assert (ov == want_ov).mean() >= 0.7


def test_nms_approximate_overlap_brackets_the_reference(emul, capi):
    """geom.cuh: overlap_approx / overlap_approx_band -- the filter that lets nms_mask_kernel decide IoU > thresh without the
    clip.  For every pair the filter accepts (overlap_approx_usable: positive sizes, relative heading not within 1e-3 rad of
    a multiple of 90 degrees) the reference's overlap (oracle, GPU dialect) must lie in [approx - slack, approx + band + slack]: proposal clusters (cars, pedestrians-sized, mixed), far coordinates, the
    adversarial set (identical boxes, shared edges, corners at MARGIN +- ulp, 90-degree turns)."""
    import math
    sets = []
    for seed in range(3):
        p, _ = synth.proposals(500, 8, seed)
        sets.append(p.numpy())
    small = synth.proposals(400, 6, 7)[0].numpy().copy()
    small[:, 3:5] *= np.array([0.2, 0.4], dtype=np.float32)          # 0.8 x 0.6 m boxes: the MARGIN band is large relative to them
    small[:, :2] = small[:, :2] * 0.25 + 30.0
    sets.append(small)
    far = synth.proposals(300, 5, 9)[0].numpy().copy(); far[:, :2] += 5000.0
    sets.append(far)
    base = np.array([10.0, 5.0, -1.0, 3.9, 1.6, 1.5, 0.3], dtype=np.float32)
    rows = [base.copy()]
    for dxy in (0.0, 1e-3, 0.00999, 0.01, 0.01001, 0.02, 0.5, 1.6, 1.61, 3.9, 3.91):
        for ang in (0.0, 0.3, 0.3 + math.pi / 2, 0.3 + math.pi, 1.57, -2.8):
            for sgn in (0, 1):
                b = base.copy()
                b[0] += (dxy * math.cos(0.3)) if sgn == 0 else (-dxy * math.sin(0.3))
                b[1] += (dxy * math.sin(0.3)) if sgn == 0 else (dxy * math.cos(0.3))
                b[6] = ang
                rows.append(b)
    sets.append(np.stack(rows).astype(np.float32))
    checked = 0
    for boxes in sets:
        n = boxes.shape[0]
        ref = capi.boxes_overlap_bev(boxes, boxes, dialect=capi.GPU)
        i, j = np.nonzero(np.triu(np.ones((n, n), dtype=bool), 1))
        a, b = np.ascontiguousarray(boxes[i]), np.ascontiguousarray(boxes[j])
        ta, tb = trig4(a), trig4(b)
        m = a.shape[0]
        approx, slack, band = (np.empty(m, np.float32) for _ in range(3))
        usable = np.empty(m, np.int32)
        emul.emul_overlap_approx(a.ctypes.data, ta.ctypes.data, b.ctypes.data, tb.ctypes.data, m, approx.ctypes.data, slack.ctypes.data, band.ctypes.data, usable.ctypes.data)
        u = usable == 1
        a, b, approx, slack, band, r = a[u], b[u], approx[u], slack[u], band[u], ref[i, j][u]
        m = int(u.sum())
        assert np.isfinite(approx).all()
        lo_ok, hi_ok = r >= approx - slack, r <= approx + band + slack
        assert lo_ok.all() and hi_ok.all(), (int((~lo_ok).sum()), int((~hi_ok).sum()), float((approx - slack - r).max()), float((r - approx - band - slack).max()))
        # and the filter is worth having: the interval is narrow next to the overlaps it has to classify
        big = r > 0.3 * np.minimum(a[:, 3] * a[:, 4], b[:, 3] * b[:, 4])
        if big.any():
            assert np.median((band + 2 * slack)[big] / r[big]) < 0.2
        checked += m
    assert checked > 350000
