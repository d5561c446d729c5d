set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cg_reduce_kernel|confine_kernel|refine_path_kernel" -c 4 -o gpurun_out/r01_rediscretize python tools/rediscretize_bench.py --reps 1 --only C2 > gpurun_out/ncu_rd.log 2>&1
tail -5 gpurun_out/ncu_rd.log
