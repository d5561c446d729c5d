#!/bin/bash
mkdir -p gpurun_out/r02
T=${1:-v15}
python -m pytest tests -m gpu -q > gpurun_out/r02/gpu_tests_$T.log 2>&1
tail -3 gpurun_out/r02/gpu_tests_$T.log
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02/bench_ref_$T.json 2> gpurun_out/r02/bench_ref_$T.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r02/bench_$T.json 2> gpurun_out/r02/bench_$T.err
echo "rc=$?" >> gpurun_out/r02/bench_$T.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mc_sim_kernel --launch-skip 3 -c 1 -o gpurun_out/r02/${T}_w1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r02/ncu_$T.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r02/bench_under_ncu_$T.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/r02/bench_$T.json')); print('%.4g %.4g'%(d['value'],d['e2e']['value']), d['roofline']['frac'], d.get('equal_work'))
"
