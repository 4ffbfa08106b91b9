#!/bin/bash
# 8-GPU visit: training step (global batch 64, 8 per GPU) and sampling (one batch of 16 sharded) through bench.py under torchrun
TAG=${1:-n8}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
nvidia-smi -L > ${L}_smi.log 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload train --steps 10 --warmup 3 > ${L}_bench_train.json 2> ${L}_bench_train.err; echo "rc=$?" >> ${L}_bench_train.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 30 --warmup 5 > ${L}_bench.json 2> ${L}_bench.err; echo "rc=$?" >> ${L}_bench.err
grep -c "comm 0x" ${L}_bench_train.err; tail -n 2 ${L}_bench_train.err; head -c 1500 ${L}_bench_train.json
tail -n 2 ${L}_bench.err; head -c 700 ${L}_bench.json
