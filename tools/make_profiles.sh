#!/bin/bash
# Round profiles (run under gpurun on 1 GPU): launch list of the bench command + ncu --set full captures
# of the dominant kernels.  The .ncu-rep files come back in gpurun_out/; tools/ncu_summary.py and
# tools/ncu_lines.py turn them into the text summaries committed under profiles/.
set -u
R=${1:-r01}
mkdir -p gpurun_out
# 1. every launch of the bench command with its device time (cold cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
# 2. full captures
K='regex:pib_query|iou_tile|iou_aligned|nms_mask|nms_sweep|pib_build'
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s $2 -c $3 -f -o gpurun_out/${R}_$1 python tools/prof_workloads.py $1 3 2>&1 | tail -1; }
cap iou_frames 2 1
cap iou_sparse 2 1
cap pib 4 2
cap nms 4 2
cap iou_dense 2 1
