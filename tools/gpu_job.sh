set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:mc_sim_kernel --launch-skip 2 -c 1 -o gpurun_out/r01_v14_mc_sim_step_r148 python bench.py --replicas 148 --warps 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_v14_r148.log 2>&1
tail -2 gpurun_out/ncu_full_v14_r148.log | cut -c1-200
