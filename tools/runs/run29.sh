set -x
mkdir -p gpurun_out
Z0=-DGLENET_PIB_ZSLABS=0
VARIANTS="a_old:$Z0,-DGLENET_PIB_BUILD_SPLIT=0,-DGLENET_PIB_PREFILL=0 b_old_split:$Z0,-DGLENET_PIB_PREFILL=0 c_old_prefill:$Z0,-DGLENET_PIB_BUILD_SPLIT=0 d_old_both:$Z0 e_new: g_new_pf1:-DGLENET_PIB_PF=1 h_new_pf0:-DGLENET_PIB_PF=0 i_new_pf0_l2pf2:-DGLENET_PIB_PF=0,-DGLENET_PIB_L2PF=2 j_new_noprefill:-DGLENET_PIB_PREFILL=0 k_old_pf1:$Z0,-DGLENET_PIB_PF=1 l_new_build_only:-DGLENET_PIB_DBG=16 m_old_ctas2:$Z0,-DGLENET_PIB_CTAS=2 n_new_ctas1:-DGLENET_PIB_CTAS=1" bash tools/pib_variants.sh > gpurun_out/pib_variants_build29.log 2>&1
tail -3 gpurun_out/pib_variants_build29.log
python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run29.log
