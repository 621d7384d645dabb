set -x
timeout 600 python -m pytest tests/test_exchange_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_exchange_2gpu_run18.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_run18.json 2> gpurun_out/bench_n2_run18.err; tail -3 gpurun_out/bench_n2_run18.err
