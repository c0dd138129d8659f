#!/bin/bash
# 2-GPU validation: torchrun bench (NCCL all-gather of the sharded outputs) + reference arm under torchrun
mkdir -p gpurun_out
N=${1:-2}
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -3 ) > gpurun_out/bench_n$N.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 ) > gpurun_out/bench_n1.log 2>&1
echo "== N=$N"; cat gpurun_out/bench_n$N.log; echo "== N=1"; cat gpurun_out/bench_n1.log
