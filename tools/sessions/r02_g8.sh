#!/bin/bash
# round-2 GPU session G8 (8 GPUs): the bench line with the strong_cfg5 sub-record at 8 ranks
set -u
O=gpurun_out
mkdir -p $O
N=${1:-8}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 > $O/g${N}_bench.json 2> $O/g${N}_bench.err ) 2> $O/g${N}_bench.time
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tests/mgpu_check.py > $O/g${N}_mgpu_check.log 2>&1
