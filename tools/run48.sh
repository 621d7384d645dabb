timeout 300 ncu --set full --clock-control none --import-source on -k regex:nms_mask_spatial\|nms_spatial -s 4 -c 2 -f -o gpurun_out/nms_sp python tools/prof_workloads.py nms 3 2>&1 | tail -1
