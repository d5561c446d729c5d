#!/bin/bash
mkdir -p gpurun_out/r02
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs --e2e-steps 1"
N="timeout 600 ncu --set full --clock-control none -k regex:mc_sim_kernel --launch-skip 3 -c 1"
$N -o gpurun_out/r02/p5_mb2_1024 $B --lib tools/_variants/mb2.so --table-slots 256 > /dev/null 2>&1
$N -o gpurun_out/r02/p5_mb2_2048 $B --lib tools/_variants/mb2.so --table-slots 256 --replicas 2048 > /dev/null 2>&1
ls -la gpurun_out/r02/p5*
