#!/bin/bash
# round-2 probe: baseline vs two warps per replica at 7 replicas per block
mkdir -p gpurun_out/r02
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --warps 1 > gpurun_out/r02/probe_w1.json 2> gpurun_out/r02/probe_w1.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --warps 2 --rpb 7 > gpurun_out/r02/probe_w2.json 2> gpurun_out/r02/probe_w2.err
python tools/phase_timers.py --lib tools/_variants/timers.so --warps 1,2 --out gpurun_out/r02/phase_timers_v14.json > gpurun_out/r02/phase_timers_v14.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mc_sim_kernel --launch-skip 2 -c 1 -o gpurun_out/r02/v14_w2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --warps 2 --rpb 7 > gpurun_out/r02/ncu_w2.log 2>&1
echo done
