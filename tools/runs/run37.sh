VARIANTS="a_z0_s8:-DGLENET_PIB_ZSLABS=0,-DGLENET_PIB_REC_STRIDE=8 b_def: c_build_only:-DGLENET_PIB_DBG=16 d_z0:-DGLENET_PIB_ZSLABS=0" bash tools/pib_variants.sh > gpurun_out/pib_variants_build37.log 2>&1
grep -i error gpurun_out/pib_variants_build37.log
python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run37.log
