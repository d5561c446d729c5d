set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/c5_exchange.py --replicas 2048 --beads 10000 --rounds 5 --sweeps 10 --out gpurun_out/c5_exchange_n2.json > gpurun_out/c5_exchange_n2.log 2>&1
tail -5 gpurun_out/c5_exchange_n2.log
