"""world_size-2 gloo test (CPU) of the host-side multi-rank logic: chain / observation
partitioning, communicator-id exchange, gathering along the chain axis."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from extensiblemcmc_jl_b200 import parallel as par


def test_shard_helpers_cover_everything_exactly_once():
    for n, w in [(4096, 8), (10, 3), (7, 8), (65536, 8), (1, 2)]:
        spans = [par.shard_chains(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2
    for n, w in [(10**9, 8), (1_000_001, 4), (5, 2), (2, 4)]:
        spans = [par.shard_obs(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2 and o2 % 2 == 0


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cid = par.exchange_comm_id(dist, make_id=lambda: bytes(range(128)))
        off, cnt = par.shard_chains(10, rank, world)
        b = par.backend_for_rank(rank, world, 0, 10, shard="chains", seed=5)
        bo = par.backend_for_rank(rank, world, 0, 10, shard="obs", comm_id=cid)
        local = np.arange(off, off + cnt, dtype=np.float64)[None, :] * np.ones((3, 1))
        full = par.gather_chain_axis(dist, local)
        q.put((rank, cid, (b.chain_offset, b.n_chains, b.shard_mode), (bo.n_chains, bo.shard_mode, bo.comm_id == cid),
               None if full is None else full.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_rendezvous_and_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] == bytes(range(128)) for r in res)
    assert res[0][2] == (0, 5, "chains") and res[1][2] == (5, 5, "chains")
    assert res[0][3] == (10, "obs", True) and res[1][3] == (10, "obs", True)
    assert res[0][4] == (np.arange(10.0)[None, :] * np.ones((3, 1))).tolist() and res[1][4] is None
