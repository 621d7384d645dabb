set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nms" 2>&1 | tail -4
timeout 600 python tools/variants_time.py 2>&1 | tee gpurun_out/variants_run24.log
