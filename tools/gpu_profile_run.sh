#!/bin/bash
# ncu evidence for profiles/ (default per-layer path; SR3_MEGA=1 variants for the step kernel):
#  (1) launch list of the benchmark step, (2) time / DRAM / tensor-op counters of every launch of one step (-> roofline.traffic, tensor-pipe %),
#  (3) --set full with source of the epilogue-bound tile shape, (4) the same counters for one step-kernel launch.
# usage: tools/gpu_profile_run.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${L}_launches.csv python tools/profile_one_step.py 4 16 > ${L}_ncu_a.log 2>&1
timeout 1500 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file ${L}_step_metrics.csv python tools/profile_one_step.py 4 16 > ${L}_ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tile_kernel -s 4 -c 1 -o ${L}_hi64_full python -c "
import sys; sys.path.insert(0,'.')
import torch, sr3_b200
from sr3_b200 import _native
torch.zeros(1).cuda()
print(_native.bench_conv(16,128,128,64,64,reps=3))" > ${L}_ncu_d.log 2>&1
SR3_MEGA=1 timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file ${L}_mega_metrics.csv python tools/profile_one_step.py 5 16 > ${L}_ncu_e.log 2>&1
ls -la gpurun_out | tail -8
tail -3 ${L}_ncu_a.log ${L}_ncu_b.log ${L}_ncu_d.log ${L}_ncu_e.log
