set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_r01_final.log 2>&1
tail -6 gpurun_out/gpu_tests_r01_final.log
