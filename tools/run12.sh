set -x
timeout 900 python -m pytest tests/test_rotate_iou_crop_gpu.py -m gpu -q 2>&1 | tail -30 | tee gpurun_out/pytest_rotate_crop_run12.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_run12.log
