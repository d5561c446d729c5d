#!/bin/bash
# table capacity with a variant library: tools/_variants/probe12.sh <lib> 768 800 832 ...
L=$1; shift
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra-legs --e2e-steps 1 --lib tools/_variants/$L.so"
for s in "$@"; do
  $B --table-slots $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($s, d['value'], d['ms_per_step'])"
done
