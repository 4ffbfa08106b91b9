mkdir -p gpurun_out
timeout 600 python tools/gpu_diag.py > gpurun_out/diag10.log 2>&1; grep -E "SUMMARY|eps rel" gpurun_out/diag10.log | tail -2
grep -iE "error|timeout|Traceback" gpurun_out/diag10.log | head -5
for v in 0 1; do
  if [ $v = 1 ]; then export SR3_NO_CSK=1; else unset SR3_NO_CSK; fi
  timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NO_CSK=$v', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
  timeout 300 python tools/bench_configs.py 2>&1 | tail -4 | cut -c1-100
done
unset SR3_NO_CSK
timeout 900 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -2
