#!/bin/bash
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-legs --e2e-steps 1"
for v in std head bb std head; do
  if [ $v = std ]; then L=""; else L="--lib tools/_variants/$v.so"; fi
  $B $L 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'])"
done
