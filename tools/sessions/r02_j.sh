#!/bin/bash
# chains-sweep experiment: chains per thread (R) x CTAs per SM
set -u
O=gpurun_out; mkdir -p $O
: > $O/j_rxctas.txt
for cfg in "18 8" "18 9" "14 8" "14 12" "14 16" "12 16" "12 24"; do
  set -- $cfg
  echo "variant=$1 ctas=$2" >> $O/j_rxctas.txt
  SWEEP_VARIANT=$1 EXTMCMC_CHAINS_CTAS=$2 timeout 120 python tools/prof_block.py cfg2 100 1 >> $O/j_rxctas.txt 2>&1
done
