set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_run9.log
timeout 600 python tools/variants_time.py 2>&1 | tee gpurun_out/variants_run9.log
timeout 900 python bench.py --steps 20 > gpurun_out/bench_run9.json 2> gpurun_out/bench_run9.err; tail -3 gpurun_out/bench_run9.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_run9.json 2>&1
