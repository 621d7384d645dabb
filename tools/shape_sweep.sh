#!/bin/bash
# Developer aid: build iou_tile_kernel with different CTA shapes (see the GLENET_IOU_* macros in csrc/iou.cu) into
# glenet_b200/lib/libglenet_geom_shape_<name>.so; tools/shape_sweep.py times them on the bench's anchor sweep.
set -e
cd "$(dirname "$0")/../glenet_b200/csrc"
build() { # name threads tr_max ctas qcap zbytes
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I ../../include -shared \
       -DGLENET_IOU_THREADS=$2 -DGLENET_IOU_TR_MAX=$3 -DGLENET_IOU_CTAS=$4 -DGLENET_IOU_QCAP=$5 -DGLENET_IOU_ZBYTES=$6 \
       iou.cu iou3d_v1.cu nms.cu pib.cu host.cpp -o ../lib/libglenet_geom_shape_$1.so 2>&1 | grep -v "warning\|nms.cu\|\^\|detected\|^$\|Remark" || true
  echo "built $1"
}
VARIANTS=${VARIANTS:-"base:256:384:4:512:4096 t128:128:192:7:256:2048 t192:192:288:5:384:4096 t160:160:224:6:320:2048 t512:512:512:2:1024:4096"}
for v in $VARIANTS; do IFS=: read n t r c q z <<< "$v"; build $n $t $r $c $q $z & done
wait
