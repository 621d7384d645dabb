set -x
mkdir -p gpurun_out
P1=-DGLENET_PIB_PF=1
VARIANTS="a_z0_s8:-DGLENET_PIB_ZSLABS=0,-DGLENET_PIB_REC_STRIDE=8 b_pf1:$P1 c_pf1_l2pf2:$P1,-DGLENET_PIB_L2PF=2 d_pf1_l2pf5:$P1,-DGLENET_PIB_L2PF=5 e_pf1_t384:$P1,-DGLENET_PIB_THREADS=384 f_pf1_t320:$P1,-DGLENET_PIB_THREADS=320 g_pf1_t512:$P1,-DGLENET_PIB_THREADS=512 h_pf1_chunk4096:$P1,-DGLENET_PIB_CHUNK=4096 i_pf1_t192_c3:$P1,-DGLENET_PIB_THREADS=192,-DGLENET_PIB_CTAS=3,-DGLENET_PIB_REC_STRIDE=8 j_z0_pf1:-DGLENET_PIB_ZSLABS=0,$P1 k_pf0:-DGLENET_PIB_PF=0" bash tools/pib_variants.sh > gpurun_out/pib_variants_build33.log 2>&1
grep -i "error" gpurun_out/pib_variants_build33.log | head
python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run33.log
