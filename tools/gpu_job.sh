set -x
mkdir -p gpurun_out
timeout 900 python bench.py --beads 400000 --replicas 148 --sweeps 5 --steps 4 --warmup 3 --ref-warm 100 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_c4_148x400k.json 2> gpurun_out/bench_c4.err
cat gpurun_out/bench_c4_148x400k.json | cut -c1-1500; tail -3 gpurun_out/bench_c4.err
timeout 600 python bench.py --beads 1000 --replicas 1024 --sweeps 200 --steps 5 --no-cpu-baseline > gpurun_out/bench_c1_1024x1000.json 2> gpurun_out/bench_c1.err
cat gpurun_out/bench_c1_1024x1000.json | cut -c1-1500
