#!/bin/bash
# usage: tools/gpu_run_c.sh <tag>   -- GPU tests, multi-stream experiment, conv sweeps
TAG=${1:-c}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
timeout 1200 python -m pytest tests -m gpu -x -q > ${L}_pytest.log 2>&1; echo "rc=$?" >> ${L}_pytest.log
TAG=$TAG SR3_CLUSTER_DEBUG=1 timeout 900 python tools/gpu_streams_check.py 16 8 4 2 > ${L}_streams.log 2>&1; echo "rc=$?" >> ${L}_streams.log
timeout 600 python tools/gpu_splitk_sweep.py > ${L}_splitk.log 2>&1; echo "rc=$?" >> ${L}_splitk.log
SR3_NO_CLUSTER=1 timeout 600 python tools/gpu_splitk_sweep.py > ${L}_splitk_nocluster.log 2>&1; echo "rc=$?" >> ${L}_splitk_nocluster.log
timeout 600 python tools/gpu_conv_sweep.py > ${L}_convsweep.log 2>&1; echo "rc=$?" >> ${L}_convsweep.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-secondary > ${L}_bench.json 2> ${L}_bench.err; echo "rc=$?" >> ${L}_bench.err
tail -n 8 ${L}_pytest.log
cat ${L}_streams.log | tail -30
head -c 600 ${L}_bench.json
