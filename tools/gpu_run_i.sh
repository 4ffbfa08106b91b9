#!/bin/bash
TAG=${1:-i}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
timeout 300 python tools/gpu_train_check.py tiny l2 > ${L}_train_tiny_l2.log 2>&1; echo "rc=$?" >> ${L}_train_tiny_l2.log
timeout 300 python tools/gpu_train_check.py three l2 > ${L}_train_three_l2.log 2>&1; echo "rc=$?" >> ${L}_train_three_l2.log
TAG=$TAG timeout 600 python tools/gpu_train_bench.py 8 5 > ${L}_train_bench.log 2>&1
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_unet.py -m gpu -x -q > ${L}_pytest.log 2>&1; echo "rc=$?" >> ${L}_pytest.log
tail -n 3 ${L}_train_tiny_l2.log ${L}_train_three_l2.log
tail -n 5 ${L}_train_bench.log
tail -n 4 ${L}_pytest.log
