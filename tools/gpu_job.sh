set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/gpu_tests_r01_final.log 2>&1
tail -15 gpurun_out/gpu_tests_r01_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err
cat gpurun_out/bench_r01_final.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_final_ref.json 2> gpurun_out/bench_r01_final_ref.err
cat gpurun_out/bench_r01_final_ref.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_v14_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_v14.log 2>&1
