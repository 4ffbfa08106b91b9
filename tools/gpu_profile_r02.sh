#!/bin/bash
# ncu evidence of round 2 for profiles/: usage tools/gpu_profile_r02.sh <tag>
TAG=${1:-p}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
ncu --query-metrics 2>/dev/null | grep -iE "tensor|umma|utc|tmem" > ${L}_query_metrics.txt
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_uniform.sum"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${L}_launches.csv python tools/profile_one_step.py 4 16 > ${L}_ncu_a.log 2>&1
timeout 1200 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file ${L}_step_metrics.csv python tools/profile_one_step.py 4 16 > ${L}_ncu_b.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file ${L}_train_launches.csv python tools/profile_one_train_step.py 8 > ${L}_ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 40 -c 1 --profile-from-start off -o ${L}_wgrad_full python tools/profile_one_train_step.py 8 > ${L}_ncu_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tile_kernel -s 4 -c 1 -o ${L}_hi64_full python -c "
import sys; sys.path.insert(0,'.')
import torch, sr3_b200
from sr3_b200 import _native
torch.zeros(1).cuda()
print(_native.bench_conv(16,128,128,64,64,reps=3))" > ${L}_ncu_e.log 2>&1
ls -la gpurun_out | tail -12
tail -2 ${L}_ncu_a.log ${L}_ncu_b.log ${L}_ncu_c.log ${L}_ncu_d.log ${L}_ncu_e.log
wc -l ${L}_query_metrics.txt
