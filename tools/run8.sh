set -x
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
tools/cuda/bin/ffma_peak 0
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_run8.log
timeout 600 python tools/variants_time.py 2>&1 | tee gpurun_out/variants_run8.log
timeout 300 python tools/exchange_probe.py 2>&1 | tee gpurun_out/exchange_probe_run8.json
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
