set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_rediscretize.py -m gpu -x -q ) > gpurun_out/gpu_tests_rd.log 2>&1
tail -4 gpurun_out/gpu_tests_rd.log
timeout 600 python tools/rediscretize_bench.py --out gpurun_out/rediscretize_bench.json > gpurun_out/rd_bench.log 2>&1
