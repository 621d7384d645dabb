#!/bin/bash
# ncu captures of the hot kernels (run under gpurun, 1 GPU).  Reports land in gpurun_out/.
set -u
K='regex:pib_query|iou_tile|iou_aligned|nms_mask|nms_sweep|pib_build'
run() { # workload skip count
  timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s $2 -c $3 -f -o gpurun_out/prof_$1 python tools/prof_workloads.py $1 3 2>&1 | tail -1
}
run pib 4 2
run iou_sparse 2 1
run iou_dense 2 1
run iou_dense_pair 2 1
run nms 4 2
