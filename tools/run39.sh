set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_run39.log
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_n1_run39.json 2> gpurun_out/bench_n1_run39.err
tail -3 gpurun_out/bench_n1_run39.err
