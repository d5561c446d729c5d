#!/bin/bash
mkdir -p gpurun_out/r02
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02/bench_ref_v14s.json 2> gpurun_out/r02/bench_ref_v14s.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r02/bench_v14s.json 2> gpurun_out/r02/bench_v14s.err
echo "rc=$?" >> gpurun_out/r02/bench_v14s.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mc_sim_kernel --launch-skip 3 -c 1 -o gpurun_out/r02/v14s_w1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02/ncu_v14s.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r02/gpu_tests_v14s.log 2>&1
tail -3 gpurun_out/r02/gpu_tests_v14s.log
