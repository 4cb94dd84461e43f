#!/bin/bash
# data-sum cache session (1 GPU): GPU tests, cfg 4 with and without the cache, the default bench line
set -u
O=gpurun_out; mkdir -p $O
T=${1:-r02m}
timeout 600 python -m pytest tests -m gpu -q -x > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log
tail -5 $O/${T}_pytest_gpu.log
timeout 300 python bench.py --workload cfg4 --steps 400 > $O/${T}_cfg4_cache.json 2> $O/${T}_cfg4_cache.err
EXTMCMC_DATA_CACHE=0 timeout 300 python bench.py --workload cfg4 --steps 400 > $O/${T}_cfg4_nocache.json 2> $O/${T}_cfg4_nocache.err
for e in 0 7; do EXTMCMC_PDL=$e timeout 300 python bench.py --workload cfg4 --steps 400 > $O/${T}_cfg4_cache_pdl$e.json 2>&1; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/*_cfg4_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        r=d.get("cfg4", d)
        print(f, r.get("ms_per_step"), r.get("roofline",{}).get("launches_per_iteration"), r.get("roofline",{}).get("avg_launch_ms"), r.get("accept_rate_per_update"))
    except Exception as e: print(f, "ERR", e)
PY
