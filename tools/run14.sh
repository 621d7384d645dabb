set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "points or pib" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_rotate_iou_crop_gpu.py -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_rotate_crop_run14.log
VARIANTS="a_default: c_l2pf4:-DGLENET_PIB_L2PF=4 d_l2pf2:-DGLENET_PIB_L2PF=2" bash tools/pib_variants.sh 2>&1 | tail -3
timeout 900 python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run14.log
