#!/bin/bash
TAG=${1:-g}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
timeout 1500 python -m pytest tests -m gpu -x -q > ${L}_pytest.log 2>&1; echo "rc=$?" >> ${L}_pytest.log
timeout 900 python bench.py --workload train --steps 8 --warmup 3 > ${L}_bench_train.json 2> ${L}_bench_train.err; echo "rc=$?" >> ${L}_bench_train.err
TAG=$TAG timeout 600 python tools/gpu_train_bench.py 8 5 > ${L}_train_bench.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 > ${L}_bench.json 2> ${L}_bench.err; echo "rc=$?" >> ${L}_bench.err
tail -n 6 ${L}_pytest.log
tail -n 3 ${L}_bench_train.err; head -c 1500 ${L}_bench_train.json
tail -n 4 ${L}_train_bench.log
tail -n 3 ${L}_bench.err; head -c 600 ${L}_bench.json
