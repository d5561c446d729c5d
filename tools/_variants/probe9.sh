#!/bin/bash
# lockstep of 14 replicas per block (2,048 replicas per GPU) against the default 7
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-legs --e2e-steps 1"
run() { $B "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', d['value'], d['ms_per_step'], d['launch'])"; }
run --replicas 1024 --lib tools/_variants/base.so
run --replicas 2048 --lib tools/_variants/base.so
run --replicas 2048 --lib tools/_variants/rpb14.so
run --replicas 2048 --lib tools/_variants/rpb14.so --rpb 7
run --replicas 2048 --lib tools/_variants/base.so --table-slots 256
