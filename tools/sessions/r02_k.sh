#!/bin/bash
# third-warp counters + parallel MALA proposal: tests and timings
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/k_pytest.log
for i in 1 2; do python bench.py --workload cfg4 --steps 400 > $O/k_cfg4_$i.json 2>/dev/null; done
python bench.py --workload cfg3 --steps 5 > $O/k_cfg3.json 2>/dev/null
python bench.py --workload cfg5 --cfg5-n-obs 125000000 --cfg5-iters 300 > $O/k_cfg5.json 2>/dev/null
python bench.py --steps 50 --skip-cpu --skip-hbm --skip-cfg5 --skip-extras --e2e-iters 200 > $O/k_cfg2.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $O/k_launches_cfg4_warm.csv python bench.py --workload cfg4 --steps 20 > /dev/null 2>&1
tail -3 $O/k_pytest.log
