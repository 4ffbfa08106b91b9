#!/bin/bash
TAG=${1:-m}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
timeout 1500 python -m pytest tests -m gpu -x -q > ${L}_pytest.log 2>&1; echo "rc=$?" >> ${L}_pytest.log
TAG=$TAG timeout 600 python tools/gpu_train_bench.py 8 5 > ${L}_train_bench.log 2>&1
TAG=${TAG}_nopack SR3_NO_PACK_TABLE=1 timeout 600 python tools/gpu_train_bench.py 8 5 > ${L}_train_bench_nopack.log 2>&1
tail -n 4 ${L}_pytest.log
tail -n 3 ${L}_train_bench.log; tail -n 1 ${L}_train_bench_nopack.log
