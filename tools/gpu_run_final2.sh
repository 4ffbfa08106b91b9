#!/bin/bash
TAG=${1:-f2}
mkdir -p gpurun_out
L=gpurun_out/r2_${TAG}
timeout 1500 python -m pytest tests -m gpu -q > ${L}_pytest.log 2>&1; echo "rc=$?" >> ${L}_pytest.log
timeout 300 python __graft_entry__.py smoke > ${L}_smoke.log 2>&1; echo "rc=$?" >> ${L}_smoke.log
timeout 900 python bench.py --steps 30 --warmup 5 > ${L}_bench.json 2> ${L}_bench.err; echo "rc=$?" >> ${L}_bench.err
timeout 600 python bench.py --workload train --steps 8 --warmup 3 > ${L}_bench_train.json 2> ${L}_bench_train.err; echo "rc=$?" >> ${L}_bench_train.err
tail -n 3 ${L}_pytest.log; tail -n 2 ${L}_smoke.log
tail -n 1 ${L}_bench.err; python - <<'PY'
import json
for f in ("gpurun_out/r2_f2_bench.json","gpurun_out/r2_f2_bench_train.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("cpu_baseline") or {}).get("value"), list((d.get("secondary") or {}).keys()))
        if "secondary" in d: print({k:(v.get("ms_per_step"), v.get("error")) for k,v in d["secondary"].items()})
        print("roofline", d["roofline"].get("frac"), d["roofline"].get("traffic"), d["roofline"].get("traffic_whole_step_incl_groupnorm_apply"))
    except Exception as e: print(f, "ERR", e)
PY
