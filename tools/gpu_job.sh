set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_rediscretize.py -m gpu -x -q -k "not c2_ensemble" > gpurun_out/sanitizer_memcheck_rd.log 2>&1; echo "memcheck rd rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_rediscretize.py -m gpu -x -q -k "not c2_ensemble" > gpurun_out/sanitizer_racecheck_rd.log 2>&1; echo "racecheck rd rc=$?"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity.py tests/test_api.py -m gpu -x -q -k "tw" > gpurun_out/sanitizer_memcheck_tw.log 2>&1; echo "memcheck twist rc=$?"
tail -4 gpurun_out/sanitizer_memcheck_rd.log gpurun_out/sanitizer_racecheck_rd.log gpurun_out/sanitizer_memcheck_tw.log
