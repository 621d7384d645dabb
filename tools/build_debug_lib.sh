#!/bin/bash
# Developer aid: library with -DGLENET_PHASE_TIMING (per-phase clock64 accumulators, tile-height override,
# clip / zero-fill ablation flags) used by tools/phase_timing.py and tools/tile_rows_sweep.py.
set -e
cd "$(dirname "$0")/../glenet_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DGLENET_PHASE_TIMING -Xcompiler -fPIC \
     -I ../../include -shared iou.cu iou3d_v1.cu nms.cu pib.cu host.cpp -o ../lib/libglenet_geom_dbg.so
echo built glenet_b200/lib/libglenet_geom_dbg.so
