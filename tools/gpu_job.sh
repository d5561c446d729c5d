set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_r01_final.log 2>&1
tail -6 gpurun_out/gpu_tests_r01_final.log
timeout 600 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err
cut -c1-250 gpurun_out/bench_r01_final.json
