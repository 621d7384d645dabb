set -x
timeout 600 python -m pytest tests/test_exchange_gpu.py -m gpu -x -q 2>&1 | tail -3
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n${n}_run26.json 2> gpurun_out/bench_n${n}_run26.err
  tail -2 gpurun_out/bench_n${n}_run26.err
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_n1_run26.json 2> gpurun_out/bench_n1_run26.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_ref_n8_run26.json 2>&1
