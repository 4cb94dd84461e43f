#!/bin/bash
# round-2 GPU session A (1 GPU): sanity, full bench line, cfg5 shard probes, launch lists
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/a_pytest.log
( time timeout 900 python bench.py > $O/a_bench.json 2> $O/a_bench.err ) 2> $O/a_bench.time
for pdl in 2 3; do
  EXTMCMC_PDL=$pdl timeout 300 python bench.py --workload cfg5 --cfg5-n-obs 125000000 --cfg5-iters 300 > $O/a_cfg5_125M_pdl$pdl.json 2> $O/a_cfg5_125M_pdl$pdl.err
  EXTMCMC_PDL=$pdl timeout 300 python bench.py --workload cfg5 --cfg5-n-obs 500000000 --cfg5-iters 100 > $O/a_cfg5_500M_pdl$pdl.json 2> $O/a_cfg5_500M_pdl$pdl.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/a_launches_cfg5_125M.csv \
  python bench.py --workload cfg5 --cfg5-n-obs 125000000 --cfg5-iters 20 > $O/a_ncu_cfg5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/a_launches_cfg4.csv \
  python bench.py --workload cfg4 --steps 20 > $O/a_ncu_cfg4.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/a_smi.txt
