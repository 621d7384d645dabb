set -x
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_run20.log
timeout 900 python bench.py --steps 20 > gpurun_out/bench_run20.json 2> gpurun_out/bench_run20.err; tail -2 gpurun_out/bench_run20.err
