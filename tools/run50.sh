( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_workload.py > gpurun_out/r02_memcheck.log 2>&1; tail -2 gpurun_out/r02_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_workload.py > gpurun_out/r02_racecheck.log 2>&1; tail -2 gpurun_out/r02_racecheck.log
grep "Race reported" gpurun_out/r02_racecheck.log | sed 's/.*in \([a-z_0-9]*\.cu[h]*:[0-9]*\).*/\1/' | sort | uniq -c
