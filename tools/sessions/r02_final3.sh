#!/bin/bash
# last session of round 2 (1 GPU): GPU test log and bench line of the final build; resident-CTA target of the
# chain-mapped sweep at cfg 4 (informational)
set -u
O=gpurun_out; mkdir -p $O
T=${1:-r02g}
timeout 600 python -m pytest tests -m gpu -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log
tail -3 $O/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${T}_smoke.log
( time timeout 900 python bench.py > $O/${T}_bench_1gpu.json 2> $O/${T}_bench_1gpu.err ) 2> $O/${T}_bench_1gpu.time
tail -3 $O/${T}_bench_1gpu.time
if [ "${CTAS:-0}" = 1 ]; then
for n in 4 6 8 12 16; do
  EXTMCMC_CHAINS_CTAS=$n timeout 200 python bench.py --workload cfg4 --steps 400 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('ctas $n', d['ms_per_step'], d['roofline']['avg_launch_ms'])" | tee -a $O/${T}_cfg4_ctas.txt
done
fi
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 80 -c 160 --csv --log-file $O/${T}_launches_cfg4.csv \
  python bench.py --workload cfg4 --steps 20 > $O/${T}_ncu_launches_cfg4.log 2>&1
