#!/bin/bash
mkdir -p gpurun_out/r02
nvidia-smi topo -m > gpurun_out/r02/topo.txt 2>&1
lscpu | grep -i "numa\|socket\|^CPU(s)\|model name" >> gpurun_out/r02/topo.txt 2>&1
for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -qi 0x10de $d/vendor 2>/dev/null; then echo "$d $(cat $d/numa_node) $(cat $d/class)"; fi; done >> gpurun_out/r02/topo.txt 2>&1
nproc >> gpurun_out/r02/topo.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02/bench_n8_numa.json 2> gpurun_out/r02/bench_n8_numa.err
tail -c 3000 gpurun_out/r02/bench_n8_numa.json
