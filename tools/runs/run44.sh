set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_exchange_gpu.py -m gpu -x -q 2>&1 | tail -3
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n${n}_run44.json 2> gpurun_out/bench_n${n}_run44.err
  tail -2 gpurun_out/bench_n${n}_run44.err
done
