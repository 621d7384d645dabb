set -x
mkdir -p gpurun_out
VARIANTS="a_old:-DGLENET_PIB_ZSLABS=0,-DGLENET_PIB_BUILD_SPLIT=0 b_z8: c_hyb:-DGLENET_PIB_ZSLABS=-1 d_z16:-DGLENET_PIB_ZSLABS=16 e_z8_pf1:-DGLENET_PIB_PF=1 f_hyb_pf1:-DGLENET_PIB_ZSLABS=-1,-DGLENET_PIB_PF=1 h_z8_l2pf4:-DGLENET_PIB_L2PF=4 i_z8_l2pf2:-DGLENET_PIB_L2PF=2 j_z8_chunk4096:-DGLENET_PIB_CHUNK=4096 k_z8_build_only:-DGLENET_PIB_DBG=16 l_hyb_build_only:-DGLENET_PIB_ZSLABS=-1,-DGLENET_PIB_DBG=16" bash tools/pib_variants.sh > gpurun_out/pib_variants_build30.log 2>&1
grep -i "error\|warning" gpurun_out/pib_variants_build30.log | head
python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run30.log
