#!/bin/bash
# usage: tools/gpu_run_d.sh <tag>   -- training-row bring-up: gradient check tool (tiny / three levels), then the GPU tests
TAG=${1:-d}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
timeout 300 python tools/gpu_train_check.py tiny l2 > ${L}_train_tiny_l2.log 2>&1; echo "rc=$?" >> ${L}_train_tiny_l2.log
timeout 300 python tools/gpu_train_check.py three l2 > ${L}_train_three_l2.log 2>&1; echo "rc=$?" >> ${L}_train_three_l2.log
timeout 300 python tools/gpu_train_check.py tiny l1 > ${L}_train_tiny_l1.log 2>&1; echo "rc=$?" >> ${L}_train_tiny_l1.log
timeout 900 python -m pytest tests/test_gpu_train.py -q > ${L}_pytest_train.log 2>&1; echo "rc=$?" >> ${L}_pytest_train.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_train.py > ${L}_pytest.log 2>&1; echo "rc=$?" >> ${L}_pytest.log
tail -n 50 ${L}_train_tiny_l2.log
tail -n 5 ${L}_train_three_l2.log
tail -n 5 ${L}_train_tiny_l1.log
tail -n 15 ${L}_pytest_train.log
tail -n 5 ${L}_pytest.log
