#!/bin/bash
# usage: tools/gpu_run_e.sh <tag>   -- training: tests (incl. reference wrapper), full-config training step timing
TAG=${1:-e}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_reference_wrapper.py tests/test_gpu_multi.py -m gpu -q > ${L}_pytest_train.log 2>&1; echo "rc=$?" >> ${L}_pytest_train.log
TAG=$TAG timeout 900 python tools/gpu_train_bench.py 8 5 > ${L}_train_bench.log 2>&1; echo "rc=$?" >> ${L}_train_bench.log
tail -n 25 ${L}_pytest_train.log
tail -n 12 ${L}_train_bench.log
