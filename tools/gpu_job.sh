set -x
mkdir -p gpurun_out
for slots in 256 384 512 640 768; do
timeout 600 python bench.py --table-slots $slots --steps 6 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_slots_$slots.json 2> gpurun_out/bench_slots.err
python -c "
import json; b=json.loads(open('gpurun_out/bench_slots_$slots.json').read()); print('slots=$slots', round(b['value']/1e6,1), 'M/s', b['config']['table_slots'], b['config']['replicas_per_block'])"
done
