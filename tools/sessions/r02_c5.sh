#!/bin/bash
# cfg5 only (strong scaling, N ranks): both exchange modes
set -u
O=gpurun_out; mkdir -p $O
N=${1:-8}; TAG=${2:-x}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 bench.py --workload cfg5 --cfg5-iters 1000 > $O/c5_${N}_$TAG.json 2> $O/c5_${N}_$TAG.err
