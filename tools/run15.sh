set -x
timeout 900 python -m pytest tests/test_rotate_iou_crop_gpu.py -m gpu -q --tb=short -k "oracle_contraction or cvae" > gpurun_out/pytest_rotate_crop_run15.log 2>&1
tail -5 gpurun_out/pytest_rotate_crop_run15.log
python - <<'PY' > gpurun_out/cvae_debug_run15.log 2>&1
import numpy as np, torch, os
from glenet_b200 import cvae_eval_utils as C
g = np.load("tests/golden/cvae_iou3d_golden.npz")
got = C.iou3d(torch.from_numpy(g["gboxes"]).cuda(), torch.from_numpy(g["qboxes"]).cuda()).cpu().numpy()
want = g["ious"]
d = np.abs(got - want)
bad = np.where(~(d <= 1e-5))[0]
print("n bad", len(bad), "max", np.nanmax(d))
for i in bad[:20]:
    print(i, got[i], want[i], g["gboxes"][i], g["qboxes"][i])
PY
cat gpurun_out/cvae_debug_run15.log | tail -30
VARIANTS="a_default:" bash tools/pib_variants.sh 2>&1 | tail -1
timeout 900 python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run15.log
