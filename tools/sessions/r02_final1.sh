#!/bin/bash
# round-2 evidence session (1 GPU): tests, the bench line of both arms, launch list, ncu --set full captures
set -u
O=gpurun_out; mkdir -p $O
T=${1:-r02}
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log
( time timeout 900 python bench.py > $O/${T}_bench_1gpu.json 2> $O/${T}_bench_1gpu.err ) 2> $O/${T}_bench_1gpu.time
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_bench_reference_arm.json 2> $O/${T}_bench_reference_arm.err
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-hbm --skip-cfg5 --skip-extras --e2e-iters 20 > $O/${T}_plain_short.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 5 --warmup 3 --skip-cpu --skip-hbm --skip-cfg5 --skip-extras --e2e-iters 20 > $O/${T}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_gsn1d_chains -s 4 -c 1 -o $O/${T}_full_chains -f python tools/prof_block.py cfg2 6 1 > $O/${T}_ncu_chains.log 2>&1
N_OBS=268435456 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_gsn1d_obs -s 4 -c 1 -o $O/${T}_full_obs -f python tools/prof_block.py cfg5 6 1 > $O/${T}_ncu_obs.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_logistic -s 2 -c 1 -o $O/${T}_full_logistic -f python bench.py --workload cfg3 --steps 3 > $O/${T}_ncu_logistic.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'accept_kernel|mala_propose' -s 24 -c 4 -o $O/${T}_full_step_cfg4 -f python bench.py --workload cfg4 --steps 20 > $O/${T}_ncu_step_cfg4.log 2>&1
