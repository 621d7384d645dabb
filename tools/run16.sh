set -x
timeout 600 python tests/golden/make_golden_cvae_iou3d.py gpu 2>&1 | tail -3
cp gpurun_out/cvae_iou3d_gpu_golden.npz tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests/test_rotate_iou_crop_gpu.py -m gpu -q --tb=short > gpurun_out/pytest_rotate_crop_run16.log 2>&1
tail -5 gpurun_out/pytest_rotate_crop_run16.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_run16.log
