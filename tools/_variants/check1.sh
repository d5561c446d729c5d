#!/bin/bash
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q > gpurun_out/r02/gpu_tests_head.log 2>&1; tail -3 gpurun_out/r02/gpu_tests_head.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_philox_parity.py tests/test_edge_cases.py -m gpu -x -q > gpurun_out/r02/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_philox_parity.py -m gpu -x -q -k "hp1 or dense or wide" > gpurun_out/r02/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02/racecheck.log
