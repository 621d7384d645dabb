set -x
mkdir -p gpurun_out
VARIANTS="a_z0_s8:-DGLENET_PIB_ZSLABS=0,-DGLENET_PIB_REC_STRIDE=8 b_def: c_build_only:-DGLENET_PIB_DBG=16 d_build_noscan:-DGLENET_PIB_DBG=20 e_build_noscan_nocoarse:-DGLENET_PIB_DBG=22 f_build_noscan_noraster:-DGLENET_PIB_DBG=21 g_build_nothing:-DGLENET_PIB_DBG=31 h_noscan_query:-DGLENET_PIB_DBG=4" bash tools/pib_variants.sh > gpurun_out/pib_variants_build34.log 2>&1
grep -i "error" gpurun_out/pib_variants_build34.log | head
python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run34.log
