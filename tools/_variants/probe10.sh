#!/bin/bash
# table capacity at the stationary working point: tools/_variants/probe10.sh 768 704 ...
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra-legs --e2e-steps 1"
for s in "$@"; do
  $B --table-slots $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($s, d['value'], d['ms_per_step'], d['launch'])"
done
