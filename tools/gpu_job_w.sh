set -x
mkdir -p gpurun_out
for cfg in "148 1" "148 2" "296 1" "296 2" "592 1" "592 2"; do set -- $cfg
timeout 900 python bench.py --beads 100000 --replicas $1 --warps $2 --sweeps 10 --steps 4 --warmup 3 --ref-warm 200 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_w_$1_$2.json 2> gpurun_out/bench_w.err
python -c "
import json; b=json.loads(open('gpurun_out/bench_w_$1_$2.json').read()); print('R=$1 W=$2', round(b['value']/1e6,1), 'M/s', b['config']['replicas_per_block'], b['config']['table_slots'], b['config']['warps_per_replica'])"
done
