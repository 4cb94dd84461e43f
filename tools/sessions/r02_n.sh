#!/bin/bash
# which switch moves the cfg 4 accept rates: data-sum cache x PDL mask x deferral
set -u
O=gpurun_out; mkdir -p $O
T=${1:-r02n}
timeout 600 python -m pytest tests -m gpu -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log
tail -5 $O/${T}_pytest_gpu.log
for c in 0 1; do for p in 0 6; do for d in 0 1; do
  EXTMCMC_DATA_CACHE=$c EXTMCMC_PDL=$p EXTMCMC_DEFER=$d timeout 300 python bench.py --workload cfg4 --steps 100 > $O/${T}_cfg4_c${c}_p${p}_d${d}.json 2>&1
done; done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02n_cfg4_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        r=d.get("cfg4", d)
        print(f, r.get("ms_per_step"), r.get("accept_rate_per_update"))
    except Exception as e: print(f, "ERR", e)
PY
