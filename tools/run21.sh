set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nms_mask -s 4 -c 1 -f -o gpurun_out/prof_nms_mask_r02f python tools/prof_workloads.py nms 3 2>&1 | tail -1
