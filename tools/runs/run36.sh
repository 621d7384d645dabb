timeout 900 python -m pytest tests -m gpu -x -q -k "points_in_boxes or pib or crop or smoke" 2>&1 | tail -4
VARIANTS="a_z0_s8:-DGLENET_PIB_ZSLABS=0,-DGLENET_PIB_REC_STRIDE=8 b_def:" bash tools/pib_variants.sh > gpurun_out/pib_variants_build36.log 2>&1
python tools/pib_variants.py 2>&1 | tee gpurun_out/pib_variants_run36.log
