#!/bin/bash
# round-2 GPU session B (1 GPU): ncu --set full of the step kernels (cfg4, cfg5 shard) with source
set -u
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'accept_kernel|mala_propose' -s 24 -c 8 -o $O/b_step_cfg4 -f \
  python bench.py --workload cfg4 --steps 20 > $O/b_ncu_cfg4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'accept_kernel' -s 20 -c 2 -o $O/b_step_cfg5 -f \
  python bench.py --workload cfg5 --cfg5-n-obs 125000000 --cfg5-iters 20 > $O/b_ncu_cfg5.log 2>&1
ls -la $O/b_*
