set -x
mkdir -p gpurun_out
S8=-DGLENET_PIB_REC_STRIDE=8
VARIANTS="a_z0_s8:-DGLENET_PIB_ZSLABS=0,$S8 b_z16: c_z16_s8:$S8 d_z0_s12:-DGLENET_PIB_ZSLABS=0 f_z8_s8:-DGLENET_PIB_ZSLABS=8,$S8 g_hyb_s8:-DGLENET_PIB_ZSLABS=-1,$S8 h_z16_pf1:-DGLENET_PIB_PF=1 i_z16_ctas3_s8:-DGLENET_PIB_CTAS=3,$S8" bash tools/pib_variants.sh > gpurun_out/pib_variants_build32.log 2>&1
grep -i "error" gpurun_out/pib_variants_build32.log | head
python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run32.log
