#!/bin/bash
# 2-GPU session after the data-sum cache: multi-rank parity, the 2-rank GPU test, the bench line with the
# strong_cfg5 and (new at N > 1) cfg4 sub-records; launch list of a cached cfg 4 iteration on one GPU
set -u
O=gpurun_out; mkdir -p $O
T=${1:-r02p}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py > $O/${T}_mgpu_check.log 2>&1; echo "rc=$?" >> $O/${T}_mgpu_check.log
tail -2 $O/${T}_mgpu_check.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/${T}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest_multi.log
tail -2 $O/${T}_pytest_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > $O/${T}_bench_2gpu.json 2> $O/${T}_bench_2gpu.err ) 2> $O/${T}_bench_2gpu.time
tail -3 $O/${T}_bench_2gpu.time
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 60 -c 120 --csv --log-file $O/${T}_launches_cfg4.csv \
  python bench.py --workload cfg4 --steps 20 > $O/${T}_ncu_launches_cfg4.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02p_bench_2gpu.json") if l.startswith("{")][-1])
print(d["n_gpus"], d["ms_per_step"], d["value"])
print("cfg4", {k: d["cfg4"][k] for k in ("ms_per_step","value","n_gpus","chains_total")} if d.get("cfg4") else None)
print("cfg5", {m: v["ms_per_iteration"] for m, v in d["strong_cfg5"]["modes"].items()})
PY
