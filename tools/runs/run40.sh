mkdir -p gpurun_out
rm -f glenet_b200/lib/libglenet_geom_shape_*.so
VARIANTS="base:256:384:3:512:4096 tr448:256:448:3:512:4096 tr480:256:480:3:512:4096 tr512:256:512:3:512:4096 z8192:256:384:3:512:8192 z2048:256:384:3:512:2048 q384:256:384:3:384:4096 q768:256:384:3:768:4096" bash tools/shape_sweep.sh 2>&1 | tail -12
python tools/shape_sweep.py 2>&1 | tee gpurun_out/shape_sweep_run40.log
rm -f glenet_b200/lib/libglenet_geom_shape_*.so
