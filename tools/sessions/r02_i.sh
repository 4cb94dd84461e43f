#!/bin/bash
# chains-sweep occupancy experiment: CTAs per SM
set -u
O=gpurun_out; mkdir -p $O
for n in 4 5 6 8; do
  EXTMCMC_CHAINS_CTAS=$n timeout 300 python bench.py --steps 50 --skip-cpu --skip-hbm --skip-cfg5 --skip-extras --e2e-iters 200 > $O/i_cfg2_ctas$n.json 2> $O/i_cfg2_ctas$n.err
  EXTMCMC_CHAINS_CTAS=$n timeout 300 python bench.py --workload cfg4 --steps 400 > $O/i_cfg4_ctas$n.json 2> $O/i_cfg4_ctas$n.err
done
