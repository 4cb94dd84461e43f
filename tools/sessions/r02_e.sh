#!/bin/bash
# round-2 GPU session E (1 GPU): incremental cov-coop indices; warm source-level profile of the lean step kernels
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/e_pytest.log
timeout 300 python bench.py --workload cfg4 --steps 400 > $O/e_cfg4.json 2> $O/e_cfg4.err
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'accept_kernel|mala_propose' -s 24 -c 4 -o $O/e_step_cfg4_warm -f \
  python bench.py --workload cfg4 --steps 20 > $O/e_ncu_cfg4b.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'accept_kernel' -s 20 -c 2 -o $O/e_step_cfg5_warm -f \
  python bench.py --workload cfg5 --cfg5-n-obs 125000000 --cfg5-iters 20 > $O/e_ncu_cfg5.log 2>&1
