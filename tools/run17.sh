set -x
nvidia-smi topo -m | head -14
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n${n}_run17.json 2> gpurun_out/bench_n${n}_run17.err
  tail -2 gpurun_out/bench_n${n}_run17.err
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_n1_run17.json 2> gpurun_out/bench_n1_run17.err
