set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "points_in_boxes or pib or abi or crop or smoke" 2>&1 | tail -4 | tee gpurun_out/pytest_pib_run35.log
VARIANTS="a_z0_s8:-DGLENET_PIB_ZSLABS=0,-DGLENET_PIB_REC_STRIDE=8 b_def: c_build_only:-DGLENET_PIB_DBG=16 d_z0_s12:-DGLENET_PIB_ZSLABS=0 e_z8_s8:-DGLENET_PIB_ZSLABS=8,-DGLENET_PIB_REC_STRIDE=8 f_z0_build_only:-DGLENET_PIB_ZSLABS=0,-DGLENET_PIB_DBG=16" bash tools/pib_variants.sh > gpurun_out/pib_variants_build35.log 2>&1
grep -i "error" gpurun_out/pib_variants_build35.log | head
python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run35.log
