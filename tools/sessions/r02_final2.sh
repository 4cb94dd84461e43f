#!/bin/bash
# final evidence of round 2 (1 GPU) after the data-sum cache and the exchange-free logistic phase 1:
# GPU tests, smoke, the default bench line, launch lists (cfg 2 and cfg 4), ncu --set full of the logistic sweep
# and of the cached cfg 4 step kernels
set -u
O=gpurun_out; mkdir -p $O
T=${1:-r02f}
timeout 600 python -m pytest tests -m gpu -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log
tail -3 $O/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${T}_smoke.log
( time timeout 900 python bench.py > $O/${T}_bench_1gpu.json 2> $O/${T}_bench_1gpu.err ) 2> $O/${T}_bench_1gpu.time
tail -3 $O/${T}_bench_1gpu.time
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-hbm --skip-cfg5 --skip-extras --e2e-iters 20 > $O/${T}_plain_short.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 5 --warmup 3 --skip-cpu --skip-hbm --skip-cfg5 --skip-extras --e2e-iters 20 > $O/${T}_ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 600 -c 300 --csv --log-file $O/${T}_launches_cfg4.csv \
  python bench.py --workload cfg4 --steps 20 > $O/${T}_ncu_launches_cfg4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_logistic -s 2 -c 1 -o $O/${T}_full_logistic -f python bench.py --workload cfg3 --steps 3 > $O/${T}_ncu_logistic.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'accept_kernel|mala_propose' -s 24 -c 4 -o $O/${T}_full_step_cfg4 -f python bench.py --workload cfg4 --steps 20 > $O/${T}_ncu_step_cfg4.log 2>&1
ls -la $O | tail -20
