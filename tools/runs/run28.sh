set -x
mkdir -p gpurun_out
VARIANTS="a_default: k_build_only:-DGLENET_PIB_DBG=16 l_build_noraster:-DGLENET_PIB_DBG=17 m_build_nocoarse:-DGLENET_PIB_DBG=18 n_build_nopack:-DGLENET_PIB_DBG=24 o_build_nothing:-DGLENET_PIB_DBG=27 p_noraster_query:-DGLENET_PIB_DBG=1" bash tools/pib_variants.sh > gpurun_out/pib_variants_build28.log 2>&1
python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run28.log
