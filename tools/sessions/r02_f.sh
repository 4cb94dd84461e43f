#!/bin/bash
# round-2 GPU session F (1 GPU): task split in the accept kernels
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/f_pytest.log
timeout 300 python bench.py --workload cfg5 --cfg5-n-obs 125000000 --cfg5-iters 300 > $O/f_cfg5_125M.json 2> $O/f_cfg5.err
timeout 300 python bench.py --workload cfg4 --steps 400 > $O/f_cfg4.json 2> $O/f_cfg4.err
timeout 300 python bench.py --steps 50 --skip-cpu --skip-hbm --skip-cfg5 --skip-extras --e2e-iters 200 > $O/f_cfg2.json 2> $O/f_cfg2.err
timeout 300 python tools/overhead_probe.py > $O/f_overhead.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $O/f_launches_cfg4_warm.csv \
  python bench.py --workload cfg4 --steps 20 > $O/f_ncu_cfg4.log 2>&1
