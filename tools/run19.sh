set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dur_assign_run19.csv python tools/prof_workloads.py assign 4 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dur_frames_run19.csv python tools/prof_workloads.py iou_frames 4 > /dev/null 2>&1
grep -h "iou_tile\|decode" gpurun_out/dur_assign_run19.csv gpurun_out/dur_frames_run19.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,200-260 | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:iou_tile -s 2 -c 1 -f -o gpurun_out/prof_assign_r02e python tools/prof_workloads.py assign 3 2>&1 | tail -1
