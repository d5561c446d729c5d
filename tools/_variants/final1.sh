#!/bin/bash
mkdir -p gpurun_out/r02
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02/bench_ref_final.json 2> gpurun_out/r02/bench_ref_final.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r02/bench_final.json 2> gpurun_out/r02/bench_final.err
echo "rc=$?" >> gpurun_out/r02/bench_final.err
python tools/phase_timers.py --lib tools/_variants/timers.so --warps 1 --out gpurun_out/r02/phase_timers_v16.json > gpurun_out/r02/phase_timers_v16.log 2>&1
tail -5 gpurun_out/r02/phase_timers_v16.log
python -c "
import json
d=json.load(open('gpurun_out/r02/bench_final.json')); print('%.4g %.4g'%(d['value'],d['e2e']['value']), d['roofline']['frac'], d.get('equal_work'))
d=json.load(open('gpurun_out/r02/bench_ref_final.json')); print('%.4g'%(d['value']))
"
