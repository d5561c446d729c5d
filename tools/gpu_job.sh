set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mc_sim_kernel --launch-skip 2 -c 1 -o gpurun_out/r01_v14_mc_sim_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_v14.log 2>&1
tail -4 gpurun_out/ncu_full_v14.log | cut -c1-300
