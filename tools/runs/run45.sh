set -u
R=r02
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
K='regex:pib_query|iou_tile|iou_aligned|nms_mask|nms_sweep|nms_spatial|nms_component|nms_clip_list|pib_build|rotate_iou|vnms'
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s $2 -c $3 -f -o gpurun_out/${R}_$1 python tools/prof_workloads.py $1 3 2>&1 | tail -1; }
cap iou_frames 2 1
cap iou_sparse 2 1
cap assign 2 1
cap pib128 4 2
cap nms 8 4
cap iou_dense 2 1
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_workload.py > gpurun_out/${R}_memcheck.log 2>&1; tail -3 gpurun_out/${R}_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_workload.py > gpurun_out/${R}_racecheck.log 2>&1; tail -5 gpurun_out/${R}_racecheck.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
