set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_run4.log
timeout 600 python tools/variants_time.py 2>&1 | tee gpurun_out/variants_run4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iou_aligned -s 2 -c 1 -f -o gpurun_out/prof_aligned_r02b python tools/aligned_time.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_mask -s 1 -c 1 -f -o gpurun_out/prof_nms_mask_r02b python tools/prof_workloads.py nms > /dev/null 2>&1
