#!/bin/bash
# 2-GPU visit: multi-GPU tests (sampling + training), N=2 benches
TAG=${1:-h}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
nvidia-smi -L > ${L}_smi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_train.py -m gpu -q > ${L}_pytest_multi.log 2>&1; echo "rc=$?" >> ${L}_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload train --steps 8 --warmup 3 > ${L}_bench_train_n2.json 2> ${L}_bench_train_n2.err; echo "rc=$?" >> ${L}_bench_train_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > ${L}_bench_n2.json 2> ${L}_bench_n2.err; echo "rc=$?" >> ${L}_bench_n2.err
tail -n 8 ${L}_pytest_multi.log
tail -n 2 ${L}_bench_train_n2.err; head -c 1200 ${L}_bench_train_n2.json
tail -n 2 ${L}_bench_n2.err; head -c 500 ${L}_bench_n2.json
