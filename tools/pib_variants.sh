#!/bin/bash
# Developer aid: build pib.cu alone with different GLENET_PIB_* macros into glenet_b200/lib/variants/libpib_<name>.so;
# tools/pib_variants.py times them on BASELINE config 2 and checks that every variant returns identical assignments.
set -e
cd "$(dirname "$0")/../glenet_b200/csrc"
mkdir -p ../lib/variants
rm -f ../lib/variants/libpib_*.so
build() { n=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -Xcompiler -fPIC -I ../../include -shared "$@" pib.cu -o ../lib/variants/libpib_$n.so
  echo "built $n: $*"; }
VARIANTS=${VARIANTS:-"a_old:-DGLENET_PIB_RUNS=0 b_runs: i_ctas3:-DGLENET_PIB_CTAS=3 m_ctas3_c2048:-DGLENET_PIB_CTAS=3,-DGLENET_PIB_CHUNK=2048 n_ctas3_c1024:-DGLENET_PIB_CTAS=3,-DGLENET_PIB_CHUNK=1024 o_ctas3_hdr:-DGLENET_PIB_CTAS=3,-DGLENET_PIB_SMEM_HDR=1 p_ctas3_skip:-DGLENET_PIB_CTAS=3,-DGLENET_PIB_SLOTSKIP=1 q_ctas2:-DGLENET_PIB_CTAS=2 r_ctas2_u3:-DGLENET_PIB_CTAS=2,-DGLENET_PIB_UNROLL3=1 s_ctas2_u3_c2048:-DGLENET_PIB_CTAS=2,-DGLENET_PIB_UNROLL3=1,-DGLENET_PIB_CHUNK=2048 t_ctas3_old:-DGLENET_PIB_CTAS=3,-DGLENET_PIB_RUNS=0"}
for v in $VARIANTS; do n=${v%%:*}; f=${v#*:}; build $n ${f//,/ } & done
wait
