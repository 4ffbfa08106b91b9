#!/bin/bash
# One GPU-box visit, everything needed to judge a build: step-kernel check (tiny), GPU test suite, step-kernel check (full config),
# bench line.  usage: tools/gpu_round2_run.sh <tag> [skip_pytest]
TAG=${1:-x}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${L}_smi.log 2>&1
timeout 300 python tools/gpu_mega_check.py tiny > ${L}_mega_tiny.log 2>&1; echo "rc=$?" >> ${L}_mega_tiny.log
if grep -q '"eps_bit_equal": true' ${L}_mega_tiny.log && ! grep -q '"eps_bit_equal": false' ${L}_mega_tiny.log; then
  echo "step kernel OK on tiny" > ${L}_status.log
else
  echo "step kernel FAILED on tiny" > ${L}_status.log
fi
if [ -z "$2" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > ${L}_pytest.log 2>&1; echo "rc=$?" >> ${L}_pytest.log
fi
timeout 400 python tools/gpu_mega_check.py full > ${L}_mega_full.log 2>&1; echo "rc=$?" >> ${L}_mega_full.log
timeout 900 python bench.py --steps 30 --warmup 5 --profile-out ${L}_profile.json > ${L}_bench.json 2> ${L}_bench.err; echo "rc=$?" >> ${L}_bench.err
SR3_MEGA=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-secondary > ${L}_bench_mega.json 2> ${L}_bench_mega.err; echo "rc=$?" >> ${L}_bench_mega.err
cat ${L}_status.log
tail -c 1500 ${L}_mega_tiny.log
tail -n 15 ${L}_pytest.log 2>/dev/null
tail -c 2500 ${L}_mega_full.log
tail -c 600 ${L}_bench.err
head -c 1500 ${L}_bench.json
head -c 400 ${L}_bench_mega.json
