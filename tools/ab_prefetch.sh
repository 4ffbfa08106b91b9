mkdir -p gpurun_out
for rep in 1 2; do
for v in 0 1; do
  if [ $v = 1 ]; then export SR3_NO_PREFETCH=1; else unset SR3_NO_PREFETCH; fi
  timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NO_PREFETCH=$v rep$rep', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
done; done
unset SR3_NO_PREFETCH
timeout 300 python tools/bench_configs.py 2>&1 | tail -4 | cut -c1-120
timeout 600 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -2
