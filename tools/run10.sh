set -x
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm --format=csv
nvidia-smi topo -m | head -12
timeout 600 python -m pytest tests/test_exchange_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_exchange_2gpu_run10.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_run10.json 2> gpurun_out/bench_n2_run10.err; tail -5 gpurun_out/bench_n2_run10.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2_run10.json 2>&1
