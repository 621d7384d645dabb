set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_run25.log
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from glenet_b200 import iou3d_nms_utils as I, synth
dev = torch.device("cuda:0")
fb, fs = [], []
for f in range(8):
    b, s = synth.proposals(4096, 20, 20 + f); fb.append(b); fs.append(s)
fb, fs = torch.stack(fb).to(dev), torch.stack(fs).to(dev)
def ev(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3
for thr in (0.7, 0.1, 0.01):
    print("nms batch 8x4096 thr", thr, "%.1f us" % ev(lambda: I.nms_gpu_batch(fb, fs, thr)))
PY
