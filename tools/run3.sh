set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aligned or smoke or cfg3" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_run3.log
timeout 300 python tools/aligned_time.py 2>&1 | tee gpurun_out/aligned_time_run3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/probe_launches.csv python tools/exchange_probe.py > gpurun_out/probe_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iou_aligned -s 2 -c 1 -f -o gpurun_out/prof_aligned_r02 python tools/aligned_time.py > /dev/null 2>&1
ls -la gpurun_out | tail -5
