set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_run2.log
timeout 600 python tools/exchange_probe.py 2>&1 | tee gpurun_out/exchange_probe.json
