#!/bin/bash
# batch size of the lane-parallel prepare at the stationary working point
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra-legs --e2e-steps 1"
for s in "$@"; do
  $B --batch $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch', $s, d['value'], d['ms_per_step'])"
done
