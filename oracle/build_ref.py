"""Build recipe for ``oracle/_ref`` -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Compiles the *unmodified* reference sources where they lie under
``/root/reference`` into three torch extension modules:

* ``oracle/_ref/iou3d_nms_cuda.so``       <- pcdet/ops/iou3d_nms/src/{iou3d_cpu.cpp,
                                              iou3d_nms_api.cpp, iou3d_nms.cpp, iou3d_nms_kernel.cu}
* ``oracle/_ref/roiaware_pool3d_cuda.so`` <- pcdet/ops/roiaware_pool3d/src/{roiaware_pool3d.cpp,
                                              roiaware_pool3d_kernel.cu}
* ``oracle/_ref/iou3d_cuda.so``           <- pcdet/ops/iou3d/src/{iou3d.cpp, iou3d_cpu.cpp, iou3d_kernel.cu}

(the same source lists as the reference's ``setup.py:58-76``).  Nothing is copied
into the repository: only build products land in ``oracle/_ref`` which is
git-ignored (but travels to the GPU box with ``gpurun``).

* ``oracle/_ref/loss_utils_ref.py``, ``oracle/_ref/cvae_eval_utils_ref.py`` <- pcdet/utils/loss_utils.py,
  cvae_uncertainty/eval_utils/eval_utils.py, staged unmodified for the same reason as the next item (their torch
  cos / sin must run on the GPU to reproduce what the reference computes there).
* ``oracle/_ref/rotate_iou_numba.py`` <- pcdet/datasets/kitti/kitti_object_eval_python/rotate_iou.py, staged
  unmodified: a numba-CUDA module has no build step other than the JIT at import, which needs the GPU, so the "build
  product" that can travel to the GPU box is the file itself (reference sources for the box are staged in
  a git-ignored directory).  ``oracle/ref.py:rotate_iou_numba()`` imports it there.

Host code MUST be compiled with ``-O2``: ``iou3d_nms_kernel.cu:43`` declares
``check_rect_cross`` as a non-inline ``__device__`` function, for which nvcc emits
a strong host stub that calls ``exit(1)``; at ``-O0`` the CPU path in
``iou3d_cpu.cpp:67`` binds to that stub and the process dies silently.

The device code is generated for sm_100a so that, on the B200 box, the very
same reference kernels act as the GPU-dialect oracle.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("GLENET_REFERENCE", "/root/reference")

EXTS = {
    "iou3d_nms_cuda": [
        "pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp",
        "pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp",
        "pcdet/ops/iou3d_nms/src/iou3d_nms.cpp",
        "pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu",
    ],
    "roiaware_pool3d_cuda": [
        "pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp",
        "pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu",
    ],
    # pcdet/ops/iou3d (the [x1, y1, x2, y2, ry] variant behind boxes_aligned_iou3d_gpu of the IoU-aware heads).
    # The reference's own setup.py (pcdet/ops/iou3d/setup.py:9-13) lists only iou3d.cpp + iou3d_kernel.cu, which leaves
    # the *_cpu symbols that iou3d.cpp:264-277 binds undefined; iou3d_cpu.cpp next to them defines those.
    "iou3d_cuda": [
        "pcdet/ops/iou3d/src/iou3d.cpp",
        "pcdet/ops/iou3d/src/iou3d_cpu.cpp",
        "pcdet/ops/iou3d/src/iou3d_kernel.cu",
    ],
}


STAGED = {
    "rotate_iou_numba.py": "pcdet/datasets/kitti/kitti_object_eval_python/rotate_iou.py",
    # the CVAE recall IoU: Python loops over numpy scalars behind torch cos / sin -- with CUDA tensors (as eval_utils.py:217-219
    # passes them) the trigonometry is libdevice's, and the ill-conditioned intersection formula amplifies a 1-ulp difference
    # in it to ~1e-4 of IoU, so the goldens that pin the kernel have to be made on the GPU box (make_golden_cvae_iou3d.py gpu)
    "loss_utils_ref.py": "pcdet/utils/loss_utils.py",
    "cvae_eval_utils_ref.py": "cvae_uncertainty/eval_utils/eval_utils.py",
}


def stage_python() -> None:
    os.makedirs(OUT, exist_ok=True)
    for name, src in STAGED.items():
        path = os.path.join(REF, src)
        if os.path.isfile(path):
            shutil.copy2(path, os.path.join(OUT, name))


def have_reference() -> bool:
    return all(os.path.isfile(os.path.join(REF, s)) for srcs in EXTS.values() for s in srcs)


def built() -> bool:
    return all(os.path.isfile(os.path.join(OUT, n + ".so")) for n in EXTS)


def build(force: bool = False, verbose: bool = False) -> bool:
    """Build both extensions into oracle/_ref.  Returns True when they exist afterwards."""
    if have_reference():
        stage_python()
    if built() and not force:
        return True
    if not have_reference():
        return built()
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load

    os.makedirs(OUT, exist_ok=True)
    for name, srcs in EXTS.items():
        if os.path.isfile(os.path.join(OUT, name + ".so")) and not force:
            continue
        bdir = os.path.join(OUT, "build_" + name)
        os.makedirs(bdir, exist_ok=True)
        load(
            name=name,
            sources=[os.path.join(REF, s) for s in srcs],
            extra_cflags=["-O2"],
            extra_cuda_cflags=["-O3"],  # nvcc default device opt level; fmad on, no fast-math
            build_directory=bdir,
            is_python_module=False,
            verbose=verbose,
        )
        shutil.copy2(os.path.join(bdir, name + ".so"), os.path.join(OUT, name + ".so"))
        shutil.rmtree(bdir, ignore_errors=True)
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref built:", ok)
    sys.exit(0 if ok else 1)
