#!/bin/bash
# round-2 GPU session H8 (8 GPUs): eight INDEPENDENT cfg5 shards (125 M observations, no exchange) at the
# same time -- per-GPU spread of the update-step time, to separate rank skew from exchange cost
set -u
O=gpurun_out
mkdir -p $O
for i in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$i timeout 300 python bench.py --workload cfg5 --cfg5-n-obs 125000000 --cfg5-iters 1000 > $O/h8_indep_$i.json 2> $O/h8_indep_$i.err &
done
wait
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv > $O/h8_smi.txt
