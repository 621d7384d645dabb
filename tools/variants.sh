#!/bin/bash
# Developer aid: build the library with different tuning macros into glenet_b200/lib/libglenet_geom_var_<name>.so;
# tools/variants_time.py times them on the dense workloads (cfg3 aligned IoU, cfg1 NMS, cfg2 IoU3D, cfg4 sweep).
#   VARIANTS="base: al3:-DGLENET_AL_CTAS=3 nms3:-DGLENET_NMS_CTAS=3" tools/variants.sh
set -e
cd "$(dirname "$0")/../glenet_b200/csrc"
rm -f ../lib/libglenet_geom_var_*.so
build() { # name flags...
  name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I ../../include -shared "$@" \
       iou.cu iou3d_v1.cu nms.cu pib.cu vnms.cu rotate_iou.cu crop.cu host.cpp -o ../lib/libglenet_geom_var_$name.so 2>&1 | grep -E "error" || true
  echo "built $name ($*)"
}
VARIANTS=${VARIANTS:-"base:"}
for v in $VARIANTS; do n=${v%%:*}; f=${v#*:}; build $n ${f//,/ } & done
wait
