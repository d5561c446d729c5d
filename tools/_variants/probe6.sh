#!/bin/bash
mkdir -p gpurun_out/r02
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-legs --e2e-steps 1"
$B > gpurun_out/r02/p6_std.json 2>/dev/null
$B --lib tools/_variants/nobar.so > gpurun_out/r02/p6_nobar.json 2>/dev/null
python - <<PY
import json
for n in ('std','nobar'):
    try:
        d=json.load(open('gpurun_out/r02/p6_%s.json'%n)); print(n,'%.4g'%d['value'],d['ms_per_step'],d['launch'])
    except Exception as e: print(n,'failed',e)
PY
