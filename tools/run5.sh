set -x
timeout 600 python tools/diag_aligned.py 2>&1 | tee gpurun_out/diag_aligned.log
