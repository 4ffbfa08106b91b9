#!/bin/bash
TAG=${1:-f}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
timeout 900 python -m pytest tests/test_gpu_train.py -q -s -k philox > ${L}_pytest_train.log 2>&1; echo "rc=$?" >> ${L}_pytest_train.log
TAG=$TAG timeout 900 python tools/gpu_train_bench.py 8 5 > ${L}_train_bench.log 2>&1; echo "rc=$?" >> ${L}_train_bench.log
grep -E "run-to-run|passed|failed" ${L}_pytest_train.log | head
tail -n 12 ${L}_train_bench.log
