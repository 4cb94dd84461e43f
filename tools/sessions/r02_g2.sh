#!/bin/bash
# round-2 GPU session G (2 GPUs): multi-rank parity + the bench line with the strong_cfg5 sub-record
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/g2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/g2_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > $O/g2_bench.json 2> $O/g2_bench.err ) 2> $O/g2_bench.time
