set -x
mkdir -p gpurun_out
timeout 600 python bench.py --lt 100 --no-cpu-baseline > gpurun_out/bench_twist_n1.json 2> gpurun_out/bench_twist.err
cat gpurun_out/bench_twist_n1.json | cut -c1-400; tail -3 gpurun_out/bench_twist.err
