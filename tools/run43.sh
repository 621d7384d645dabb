VARIANTS="a_def: b_pingpong:-DGLENET_PIB_PINGPONG=1 c_pingpong_l2pf2:-DGLENET_PIB_PINGPONG=1,-DGLENET_PIB_L2PF=2 d_pingpong_l2pf4:-DGLENET_PIB_PINGPONG=1,-DGLENET_PIB_L2PF=4" bash tools/pib_variants.sh > gpurun_out/pib_variants_build43.log 2>&1
grep -i error gpurun_out/pib_variants_build43.log
python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run43.log
