set -x
mkdir -p gpurun_out
K='regex:pib_query'
timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s 2 -c 1 -f -o gpurun_out/r02b_pib128 python tools/prof_workloads.py pib128 3 2>&1 | tail -1
