set -x
python -c "import numba; print(numba.__version__)"
timeout 600 python tests/golden/make_golden_rotate_iou.py gpu 2>&1 | tail -5
cp gpurun_out/rotate_iou_gpu_golden.npz tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests/test_rotate_iou_crop_gpu.py -m gpu -q 2>&1 | tail -30 | tee gpurun_out/pytest_rotate_crop_run11.log
timeout 300 python -m pytest tests/test_oracle.py -q -k rotate 2>&1 | tail -5
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_run11.log
