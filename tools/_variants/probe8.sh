#!/bin/bash
# A/B of library variants on one box: tools/_variants/probe8.sh base pair base pair
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-legs --e2e-steps 1"
for v in "$@"; do
  $B --lib tools/_variants/$v.so 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'])"
done
