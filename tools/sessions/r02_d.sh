#!/bin/bash
# round-2 GPU session D (1 GPU): lean step-kernel instantiations vs the general ones
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/d_pytest.log
for lean in 1 0; do
  EXTMCMC_LEAN=$lean timeout 300 python bench.py --workload cfg5 --cfg5-n-obs 125000000 --cfg5-iters 300 > $O/d_cfg5_125M_lean$lean.json 2> $O/d_cfg5_lean$lean.err
  EXTMCMC_LEAN=$lean timeout 300 python bench.py --workload cfg4 --steps 400 > $O/d_cfg4_lean$lean.json 2> $O/d_cfg4_lean$lean.err
  EXTMCMC_LEAN=$lean timeout 300 python bench.py --steps 50 --skip-cpu --skip-hbm --skip-cfg5 --skip-extras --e2e-iters 200 > $O/d_cfg2_lean$lean.json 2> $O/d_cfg2_lean$lean.err
  EXTMCMC_LEAN=$lean timeout 300 python tools/overhead_probe.py > $O/d_overhead_lean$lean.txt 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $O/d_launches_cfg4_warm.csv \
  python bench.py --workload cfg4 --steps 20 > $O/d_ncu_cfg4.log 2>&1
