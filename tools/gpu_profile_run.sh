#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the benchmark step (step_kernel), (2) --set full of the step kernel with source,
# (3) tensor / dram counters of the step kernel, (4) source-level capture of the stand-alone tile kernel on the epilogue-bound shape.
# usage: tools/gpu_profile_run.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file ${L}_launches.csv python tools/profile_one_step.py 6 16 > ${L}_ncu_a.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:step_kernel -s 3 -c 2 --csv --log-file ${L}_step_metrics.csv python tools/profile_one_step.py 6 16 > ${L}_ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o ${L}_step_full python tools/profile_one_step.py 5 16 > ${L}_ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tile_kernel -s 4 -c 1 -o ${L}_hi64_full python -c "
import sys; sys.path.insert(0,'.')
import torch, sr3_b200
from sr3_b200 import _native
torch.zeros(1).cuda()
print(_native.bench_conv(16,128,128,64,64,reps=3))" > ${L}_ncu_d.log 2>&1
ls -la gpurun_out | tail -8
tail -3 ${L}_ncu_b.log ${L}_ncu_c.log ${L}_ncu_d.log
