#!/bin/bash
TAG=${1:-n8b}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload train --steps 10 --warmup 3 > ${L}_bench_train.json 2> ${L}_bench_train.err; echo "rc=$?" >> ${L}_bench_train.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --workload train --steps 10 --warmup 3 > ${L}_bench_train_n4.json 2> ${L}_bench_train_n4.err; echo "rc=$?" >> ${L}_bench_train_n4.err
tail -n 2 ${L}_bench_train.err; head -c 1300 ${L}_bench_train.json; echo; head -c 300 ${L}_bench_train_n4.json
