set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
tools/cuda/bin/ffma_peak 0 | tee gpurun_out/ffma_peak.json
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_run1.log
nproc; lscpu | grep -i numa; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; python -c "
import torch
p=torch.cuda.get_device_properties(0); print(p)
print([a for a in dir(p) if 'pci' in a])
"
