"""Seeded synthetic KITTI / Waymo shaped inputs used by tests and bench.py.

Value ranges follow the reference's configs: point-cloud ranges of
tools/cfgs/dataset_configs/{kitti,waymo}_dataset.yaml, anchor sizes / rotations / bottom
heights of tools/cfgs/kitti_models/GLENet_VR.yaml:60-90, anchor layout of
pcdet/models/dense_heads/target_assigner/anchor_generator.py:17-60.
Everything is float32 and generated on the CPU with an explicit torch.Generator.
"""
from __future__ import annotations

import math

import torch

KITTI_RANGE = (0.0, -40.0, -3.0, 70.4, 40.0, 1.0)
WAYMO_RANGE = (-75.2, -75.2, -2.0, 75.2, 75.2, 4.0)


def _gen(seed):
    return torch.Generator().manual_seed(int(seed))


def _u(g, n, lo, hi):
    return torch.rand(n, generator=g, dtype=torch.float32) * (hi - lo) + lo


def kitti_boxes(n, seed=0):
    g = _gen(seed)
    return torch.stack([
        _u(g, n, 0.0, 70.4), _u(g, n, -40.0, 40.0), _u(g, n, -1.5, -0.5),
        _u(g, n, 3.5, 4.3), _u(g, n, 1.45, 1.75), _u(g, n, 1.41, 1.71),
        _u(g, n, -math.pi, math.pi)], dim=1).contiguous()


def waymo_boxes(n, seed=0):
    g = _gen(seed)
    return torch.stack([
        _u(g, n, -75.2, 75.2), _u(g, n, -75.2, 75.2), _u(g, n, 0.0, 2.0),
        _u(g, n, 4.1, 5.3), _u(g, n, 1.8, 2.4), _u(g, n, 1.4, 2.0),
        _u(g, n, -math.pi, math.pi)], dim=1).contiguous()


def anchors_kitti3():
    """211 200 x 7 anchors: grid 176(x) x 200(y), 3 classes x 2 rotations, align_center=False."""
    sizes = [[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]]
    bottoms = [-1.78, -0.6, -0.6]
    rots = [0.0, 1.57]
    nx, ny = 176, 200
    x0, y0, _, x1, y1, _ = KITTI_RANGE
    xs = torch.arange(nx, dtype=torch.float32) * ((x1 - x0) / (nx - 1)) + x0
    ys = torch.arange(ny, dtype=torch.float32) * ((y1 - y0) / (ny - 1)) + y0
    out = torch.empty((ny, nx, 3, 2, 7), dtype=torch.float32)
    for ci, (sz, zb) in enumerate(zip(sizes, bottoms)):
        for ri, r in enumerate(rots):
            out[:, :, ci, ri, 0] = xs.view(1, nx)
            out[:, :, ci, ri, 1] = ys.view(ny, 1)
            out[:, :, ci, ri, 2] = zb + sz[2] / 2
            out[:, :, ci, ri, 3] = sz[0]
            out[:, :, ci, ri, 4] = sz[1]
            out[:, :, ci, ri, 5] = sz[2]
            out[:, :, ci, ri, 6] = r
    return out.view(-1, 7).contiguous()


def points(n, boxes, rng=KITTI_RANGE, frac_in=0.05, seed=0):
    """n points uniform in ``rng``; ``frac_in`` of them resampled inside randomly chosen boxes
    (+10 % extent so that some land on or just outside the faces)."""
    g = _gen(seed + 1000)
    p = torch.stack([_u(g, n, rng[0], rng[3]), _u(g, n, rng[1], rng[4]), _u(g, n, rng[2], rng[5])], dim=1)
    k = int(n * frac_in)
    if k and boxes.shape[0]:
        bi = torch.randint(0, boxes.shape[0], (k,), generator=g)
        b = boxes[bi]
        loc = (torch.rand((k, 3), generator=g, dtype=torch.float32) - 0.5) * b[:, 3:6] * 1.1
        c, s = torch.cos(b[:, 6]), torch.sin(b[:, 6])
        x = loc[:, 0] * c - loc[:, 1] * s + b[:, 0]
        y = loc[:, 0] * s + loc[:, 1] * c + b[:, 1]
        z = loc[:, 2] + b[:, 2]
        idx = torch.randperm(n, generator=g)[:k]
        p[idx] = torch.stack([x, y, z], dim=1)
    return p.contiguous()


def proposals(n, k=20, seed=0, base=None):
    """n proposals jittered around k object centres + distinct scores."""
    g = _gen(seed + 2000)
    centres = kitti_boxes(k, seed) if base is None else base
    k = centres.shape[0]
    sigma = torch.tensor([0.3, 0.3, 0.1, 0.15, 0.08, 0.08, 0.1], dtype=torch.float32)
    which = torch.randint(0, k, (n,), generator=g)
    boxes = centres[which] + torch.randn((n, 7), generator=g, dtype=torch.float32) * sigma
    scores = (torch.randperm(n, generator=g).float() + 1.0) / n
    return boxes.contiguous(), scores.contiguous()


def cvae_samples(g_objects, r, seed=0):
    """GT = kitti_boxes(g); r samples per GT = GT + N(0, sigma).  Returns (samples (g*r, 7), gt (g, 7))."""
    g = _gen(seed + 3000)
    gt = kitti_boxes(g_objects, seed)
    sigma = torch.tensor([0.15, 0.15, 0.05, 0.1, 0.05, 0.05, 0.05], dtype=torch.float32)
    samples = gt.repeat_interleave(r, dim=0) + torch.randn((g_objects * r, 7), generator=g, dtype=torch.float32) * sigma
    return samples.contiguous(), gt


def head_pairs(n, seed=0):
    """Row-aligned (prediction, regression target) boxes as the IoU-aware heads see them
    (anchor_head_kl_label.py:417-428: decoded predictions vs decoded targets of the positive anchors):
    targets = kitti_boxes, predictions = targets + noise; a few exact copies, tiny offsets, quarter turns and
    far misses are mixed in.  Returns (pred (n, 7), target (n, 7))."""
    g = _gen(seed + 4000)
    tgt = kitti_boxes(n, seed + 7)
    sigma = torch.tensor([0.25, 0.25, 0.1, 0.15, 0.08, 0.08, 0.12], dtype=torch.float32)
    pred = tgt + torch.randn((n, 7), generator=g, dtype=torch.float32) * sigma
    k = torch.arange(n)
    pred[k % 11 == 0] = tgt[k % 11 == 0]                                  # exact copies
    m = k % 11 == 1
    pred[m] = tgt[m]; pred[m, 0] += 1e-5                                  # one margin to the side
    m = k % 11 == 2
    pred[m] = tgt[m]; pred[m, 6] += 3.14159265 / 2                        # quarter turn about the same centre
    m = k % 11 == 3
    pred[m, 0] += 6.0                                                     # miss
    return pred.contiguous(), tgt.contiguous()
