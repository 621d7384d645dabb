set -x
timeout 900 python -m pytest tests/test_rotate_iou_crop_gpu.py -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_rotate_crop_run13.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pib_query -s 2 -c 1 -f -o gpurun_out/prof_pib16_r02 python tools/prof_workloads.py pib16 3 2>&1 | tail -2
timeout 900 python bench.py --steps 20 > gpurun_out/bench_run13.json 2> gpurun_out/bench_run13.err; tail -3 gpurun_out/bench_run13.err
bash tools/pib_variants.sh 2>&1 | tail -3
timeout 900 python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run13.log
