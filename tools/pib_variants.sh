#!/bin/bash
# Developer aid: build pib.cu alone with different GLENET_PIB_* macros into glenet_b200/lib/variants/libpib_<name>.so;
# tools/pib_variants.py times them on BASELINE config 2 and checks that every variant returns identical assignments.
# VARIANTS="name:-DMACRO=v,-DMACRO2=v ..." overrides the default list (the tunables are the GLENET_PIB_* macros at the top of pib.cu).
set -e
cd "$(dirname "$0")/../glenet_b200/csrc"
mkdir -p ../lib/variants
rm -f ../lib/variants/libpib_*.so
build() { n=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -Xcompiler -fPIC -I ../../include -shared "$@" pib.cu -o ../lib/variants/libpib_$n.so
  echo "built $n: $*"; }
VARIANTS=${VARIANTS:-"a_default: b_chunks:-DGLENET_PIB_RUNS=0 c_ctas4:-DGLENET_PIB_CTAS=4 d_ctas2:-DGLENET_PIB_CTAS=2 e_no_zwindow:-DGLENET_PIB_ZWINDOW=0 f_no_l2pf:-DGLENET_PIB_L2PF=0 g_l2pf2:-DGLENET_PIB_L2PF=2 h_l2pf4:-DGLENET_PIB_L2PF=4 i_chunk4096:-DGLENET_PIB_CHUNK=4096 j_build512:-DGLENET_PIB_BUILD_THREADS=512 k_build_only:-DGLENET_PIB_DBG=16"}
for v in $VARIANTS; do n=${v%%:*}; f=${v#*:}; build $n ${f//,/ } & done
wait
