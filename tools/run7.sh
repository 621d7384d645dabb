set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_run7.log
timeout 600 python tools/variants_time.py 2>&1 | tee gpurun_out/variants_run7.log
timeout 600 python tools/tile_rows_env_sweep.py 2>&1 | tee gpurun_out/tile_rows_run7.log
timeout 300 python tools/exchange_probe.py 2>&1 | tee gpurun_out/exchange_probe_run7.json
timeout 300 python tools/diag_aligned.py 2>&1 | head -8 | tee gpurun_out/diag_aligned3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iou_aligned -s 2 -c 1 -f -o gpurun_out/prof_aligned_r02d python tools/aligned_time.py > /dev/null 2>&1
