#!/bin/bash
# logistic sweep: phase-1 variants (EXTMCMC_LOGI_PHASE1 0 / 2; the session also measured partners on neighbouring
# sub-partitions, 38.1 ms, and an mbarrier hand-over instead of the second pair barrier, 34.26 ms -- both dropped)
set -u
O=gpurun_out; mkdir -p $O
T=${1:-r02o}
for m in ${MODES:-0 1}; do
  EXTMCMC_LOGI_PHASE1=$m timeout 300 python -m pytest tests/test_gpu_logistic.py -m gpu -q -x > $O/${T}_pytest_logi_pair$m.log 2>&1; tail -2 $O/${T}_pytest_logi_pair$m.log
  for r in 1 2; do
    EXTMCMC_LOGI_PHASE1=$m timeout 300 python bench.py --workload cfg3 --steps 8 > $O/${T}_cfg3_pair${m}_$r.json 2> $O/${T}_cfg3_pair${m}_$r.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02o_cfg3_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d.get("ms_per_step"), d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
PY
