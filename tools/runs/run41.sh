mkdir -p gpurun_out
rm -f glenet_b200/lib/libglenet_geom_shape_*.so
VARIANTS="base:256:384:3:512:4096 tr416:256:416:3:512:4096 tr448:256:448:3:512:4096 tr448q640:256:448:3:640:4096 tr448q768:256:448:3:768:4096 tr464:256:464:3:512:4096 tr448q640z2k:256:448:3:640:2048 tr464z2k:256:464:3:512:2048" bash tools/shape_sweep.sh 2>&1 | grep built
python tools/shape_sweep.py 2>&1 | tee gpurun_out/shape_sweep_run41.log
rm -f glenet_b200/lib/libglenet_geom_shape_*.so
