#!/bin/bash
mkdir -p gpurun_out/r02
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra-legs --e2e-steps 1"
$B --table-slots 256 > gpurun_out/r02/p4_std_1024_c256.json 2>/dev/null
$B --lib tools/_variants/mb2.so --table-slots 256 > gpurun_out/r02/p4_mb2_1024_c256.json 2>/dev/null
$B --lib tools/_variants/mb2.so --table-slots 256 --replicas 2048 > gpurun_out/r02/p4_mb2_2048_c256.json 2>/dev/null
python - <<PY
import json
for n in ('std_1024_c256','mb2_1024_c256','mb2_2048_c256'):
    try:
        d=json.load(open('gpurun_out/r02/p4_%s.json'%n)); print(n,'%.4g'%d['value'],d['ms_per_step'],d['launch'])
    except Exception as e: print(n,'failed',e)
PY
