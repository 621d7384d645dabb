set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nms or NMS" 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_run22.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dur_nms_run22.csv python tools/prof_workloads.py nms 4 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nms_mask -s 4 -c 1 -f -o gpurun_out/prof_nms_mask_r02g python tools/prof_workloads.py nms 3 2>&1 | tail -1
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
