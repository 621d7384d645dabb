"""The C-ABI library loads and exports every symbol include/*.h declares (no GPU needed)."""
import ctypes
import glob
import os
import re
import subprocess

from conftest import ROOT


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(glenet_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("glenet_boxes_iou_bev_gpu", "glenet_boxes_overlap_bev_gpu", "glenet_boxes_iou3d_gpu", "glenet_nms_gpu",
                 "glenet_nms_normal_gpu", "glenet_points_in_boxes_gpu", "glenet_points_in_boxes_cpu_dialect",
                 "glenet_boxes_iou_bev_cpu_dialect", "glenet_last_error"):
        assert must in syms


def test_library_builds_loads_and_exports_everything():
    import glenet_b200
    lib = glenet_b200.load()
    assert os.path.isfile(glenet_b200.lib_path())
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    # and the binding table covers exactly the header
    assert sorted(glenet_b200.EXPORTS) == declared_symbols()
    assert lib.glenet_abi_version() == 13
    # mask + deferred-clip list (16-byte counter + 32 n entries) + spatial tiles (permutation, table of up to 64 + 64 groups,
    # 4 work items per group pair, counter) + component sweep (labels + count, done counter, kept bits, membership bitmaps)
    assert lib.glenet_nms_workspace_bytes(1, 4096) == (4096 * 64 * 8 + 16 + 32 * 4096 * 8 + 4096 * 4 + 128 * 8 + 128 * 129 // 2 * 4 * 4 + 16
                                                       + 528 + 16 + 64 * 8 + 128 * 64 * 8)
    assert lib.glenet_points_in_boxes_workspace_bytes(2, 200) > 2 * 200 * 32
    # workspace layout of csrc/pib.cu (pib_layout): per frame a 48-byte header, 8 floats per box, 32 N + 8192 list slots
    # (slices of crowded coarse cells), a 128 x 128 fine map of 16 z-slab bits and 4096 packed 8-byte coarse cells; every
    # section 16-byte aligned
    up = lambda v: (v + 15) // 16 * 16
    for b, n in ((1, 1), (2, 200), (128, 200), (3, 77), (5, 4096)):
        want = up(b * 48) + up(b * n * 32) + up(b * (32 * n + 8192) * 4) + up(b * 128 * 128 * 2) + up(b * 4096 * 8)
        assert lib.glenet_points_in_boxes_workspace_bytes(b, n) == want, (b, n)
    assert lib.glenet_points_in_boxes_workspace_bytes(0, 5) == 16 and lib.glenet_points_in_boxes_workspace_bytes(4, 0) == 16


def test_library_is_sm100a_only():
    import glenet_b200
    out = subprocess.run(["cuobjdump", "-lelf", glenet_b200.lib_path()], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_torch_types_in_abi():
    import re
    text = open(os.path.join(ROOT, "include", "glenet_geom.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)       # the comments cite the reference's torch-based wrappers
    assert "at::" not in code and "torch" not in code.lower() and "tensor" not in code.lower() and "#include <cuda" not in code


def test_product_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "glenet_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        assert "oracle" not in src, f"{path} references the oracle"
    for path in glob.glob(os.path.join(ROOT, "glenet_b200", "csrc", "*")):
        assert "oracle" not in open(path).read()


def test_host_trig_helpers_match_libm():
    import numpy as np
    import glenet_b200
    lib = glenet_b200.load()
    libm = ctypes.CDLL("libm.so.6")
    libm.cosf.restype = libm.sinf.restype = ctypes.c_float
    libm.cosf.argtypes = libm.sinf.argtypes = [ctypes.c_float]
    boxes = np.zeros((50, 7), dtype=np.float32)
    boxes[:, 6] = np.linspace(-7, 7, 50, dtype=np.float32)
    out4 = np.zeros((50, 4), dtype=np.float32)
    out2 = np.zeros((50, 2), dtype=np.float32)
    lib.glenet_host_trig4(boxes.ctypes.data, 50, out4.ctypes.data)
    lib.glenet_host_trig2(boxes.ctypes.data, 50, out2.ctypes.data)
    for i in range(50):
        h = float(boxes[i, 6])
        assert out4[i, 0] == np.float32(libm.cosf(h)) and out4[i, 1] == np.float32(libm.sinf(h))
        assert out4[i, 2] == np.float32(libm.cosf(-h)) and out4[i, 3] == np.float32(libm.sinf(-h))
        assert out2[i, 0] == out4[i, 2] and out2[i, 1] == out4[i, 3]
