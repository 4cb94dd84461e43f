#!/bin/bash
# round-2 GPU session C (1 GPU): L2 evict_first hint + obs-sweep PDL; warm-cache profiles of the step kernels
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c_pytest.log
for hint in 1 0; do for pdl in 6 2; do
  EXTMCMC_L2_HINT=$hint EXTMCMC_PDL=$pdl timeout 300 python bench.py --workload cfg5 --cfg5-n-obs 125000000 --cfg5-iters 300 > $O/c_cfg5_125M_h${hint}_p$pdl.json 2> $O/c_cfg5_125M_h${hint}_p$pdl.err
done; done
EXTMCMC_L2_HINT=1 timeout 300 python bench.py --workload cfg5 --cfg5-n-obs 1000000000 --cfg5-iters 100 > $O/c_cfg5_1G_h1.json 2>$O/c_cfg5_1G_h1.err
EXTMCMC_L2_HINT=0 EXTMCMC_PDL=2 timeout 300 python bench.py --workload cfg5 --cfg5-n-obs 1000000000 --cfg5-iters 100 > $O/c_cfg5_1G_h0.json 2>$O/c_cfg5_1G_h0.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file $O/c_launches_cfg5_125M_warm.csv \
  python bench.py --workload cfg5 --cfg5-n-obs 125000000 --cfg5-iters 20 > $O/c_ncu_cfg5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $O/c_launches_cfg4_warm.csv \
  python bench.py --workload cfg4 --steps 20 > $O/c_ncu_cfg4.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'accept_kernel|mala_propose' -s 24 -c 4 -o $O/c_step_cfg4_warm -f \
  python bench.py --workload cfg4 --steps 20 > $O/c_ncu_cfg4b.log 2>&1
