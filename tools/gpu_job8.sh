set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
cat gpurun_out/bench_n8.json | cut -c1-400; tail -3 gpurun_out/bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/c5_exchange.py --replicas 4096 --beads 10000 --rounds 5 --sweeps 10 --out gpurun_out/c5_exchange_n8.json > gpurun_out/c5_exchange_n8.log 2>&1
cat gpurun_out/c5_exchange_n8.json; tail -3 gpurun_out/c5_exchange_n8.log
