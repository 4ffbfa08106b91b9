#!/bin/bash
TAG=${1:-l}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
TAG=${TAG}_w1 timeout 600 python tools/gpu_train_bench.py 8 5 > ${L}_train_bench_w1.log 2>&1
TAG=${TAG}_w2 SR3_WGRAD_WAVES=2 timeout 600 python tools/gpu_train_bench.py 8 5 > ${L}_train_bench_w2.log 2>&1
TAG=${TAG}_w3 SR3_WGRAD_WAVES=3 timeout 600 python tools/gpu_train_bench.py 8 5 > ${L}_train_bench_w3.log 2>&1
timeout 300 python tools/gpu_train_check.py tiny l2 > ${L}_train_tiny_l2.log 2>&1; echo "rc=$?" >> ${L}_train_tiny_l2.log
tail -n 3 ${L}_train_bench_w1.log; tail -n 3 ${L}_train_bench_w2.log; tail -n 3 ${L}_train_bench_w3.log; tail -n 2 ${L}_train_tiny_l2.log
