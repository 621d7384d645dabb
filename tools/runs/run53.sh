mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -2 gpurun_out/r02_bench.err
