mkdir -p gpurun_out
K='regex:pib_query|iou_tile|iou_aligned|nms_mask|nms_sweep|nms_clip_list|pib_build|rotate_iou|vnms'
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s $2 -c $3 -f -o gpurun_out/r02c_$1 python tools/prof_workloads.py $1 3 2>&1 | tail -1; }
cap iou_frames 2 1
cap iou_sparse 2 1
cap nms 6 3
cap iou_dense 2 1
