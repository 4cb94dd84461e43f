"""Multi-GPU parity (chains shard, observations shard with NCCL and with the fused peer exchange,
observation-sharded logistic regression): runs tests/mgpu_check.py under torchrun when the box
has at least two GPUs; skipped on single-GPU boxes (the script is also run by hand with
`gpurun --gpus 2|8`, results in profiles/)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert rep["chains_bitexact"] and rep["obs_decisions_equal"] and rep["p2p_matches_nccl"]
    assert rep["logistic_obs_sharded_ok"] and rep["few_chains_fused_tail_ok"] and rep["obs_ll_rel_err"] < 1e-10
