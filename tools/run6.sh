set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_run6.log
timeout 600 python tools/variants_time.py 2>&1 | tee gpurun_out/variants_run6.log
timeout 300 python tools/diag_aligned.py 2>&1 | head -8 | tee gpurun_out/diag_aligned2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iou_aligned -s 2 -c 1 -f -o gpurun_out/prof_aligned_r02c python tools/aligned_time.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iou_tile -s 1 -c 1 -f -o gpurun_out/prof_assign_r02c python tools/prof_workloads.py assign > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iou_tile -s 1 -c 1 -f -o gpurun_out/prof_frames_r02c python tools/prof_workloads.py iou_frames > /dev/null 2>&1
