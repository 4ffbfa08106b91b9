#!/bin/bash
# 1-GPU final visit of the round: whole GPU suite, smoke, default bench line, training bench, ncu launch list of a training iteration
TAG=${1:-f1}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
timeout 1500 python -m pytest tests -m gpu -q > ${L}_pytest.log 2>&1; echo "rc=$?" >> ${L}_pytest.log
timeout 300 python __graft_entry__.py smoke > ${L}_smoke.log 2>&1; echo "rc=$?" >> ${L}_smoke.log
timeout 900 python bench.py --steps 30 --warmup 5 > ${L}_bench.json 2> ${L}_bench.err; echo "rc=$?" >> ${L}_bench.err
timeout 600 python bench.py --workload train --steps 8 --warmup 3 > ${L}_bench_train.json 2> ${L}_bench_train.err; echo "rc=$?" >> ${L}_bench_train.err
TAG=$TAG timeout 600 python tools/gpu_train_bench.py 8 5 > ${L}_train_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum --clock-control none --profile-from-start off --csv --log-file ${L}_train_launches.csv python tools/profile_one_train_step.py 8 > ${L}_ncu_c.log 2>&1
tail -n 4 ${L}_pytest.log; tail -n 3 ${L}_smoke.log
tail -n 2 ${L}_bench.err; head -c 400 ${L}_bench.json; echo
tail -n 2 ${L}_bench_train.err; head -c 400 ${L}_bench_train.json; echo
tail -n 3 ${L}_train_bench.log; tail -n 2 ${L}_ncu_c.log
