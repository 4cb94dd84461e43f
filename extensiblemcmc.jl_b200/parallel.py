"""One process per GPU: how chains / observations are split and how ranks rendezvous.

The reference has no multi-process code at all.  Two modes:
  * chains shard (default): rank r owns a contiguous range of GLOBAL chain ids, the
    observations are replicated, nothing is exchanged while sampling.  Because the Philox
    key space is indexed by the global chain id, results do not depend on the sharding.
  * observations shard: every rank holds all chains (state and RNG replicated, so all
    ranks take identical decisions) and a slice of the observations; the per-chain partial
    sums are all-reduced with NCCL once per update step inside libextmcmc_cuda.
torch.distributed is used for the rendezvous only (shipping the 128-byte NCCL unique id).
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from .workspaces import CUDAMCMCBackend


def env_rank_world():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
            int(os.environ.get("LOCAL_RANK", 0)))


def shard_chains(n_total, rank, world):
    """-> (offset, count) of the contiguous chain range owned by `rank`."""
    base, rem = divmod(int(n_total), int(world))
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def shard_obs(n_total, rank, world):
    """-> (first, count); boundaries fall on even indices (16-byte units of the sweep)."""
    pairs = (int(n_total) + 1) // 2
    lo = 2 * (rank * pairs // world)
    hi = min(int(n_total), 2 * ((rank + 1) * pairs // world))
    return lo, hi - lo


def nccl_unique_id():
    buf = (C.c_uint8 * 128)()
    rc = _abi.load().extmcmc_comm_unique_id(buf)
    if rc != _abi.OK:
        raise _abi.ExtMCMCError(rc, (_abi.load().extmcmc_last_error(None) or b"").decode())
    return bytes(buf)


def exchange_comm_id(dist, make_id=nccl_unique_id):
    """Rank 0 creates the communicator id, everybody receives it (any backend)."""
    ids = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    return ids[0]


def p2p_allgather_fn(dist):
    """bytes -> list of every rank's bytes (rank order), for CUDAMCMCBackend(p2p_allgather=...)."""
    def gather(b):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, b)
        return out
    return gather


def backend_for_rank(rank, world, local_rank, n_chains_total, shard="chains", comm_id=None, **kw):
    """CUDAMCMCBackend of this rank.  shard='chains': n_chains_total is split; shard='obs':
    every rank runs all n_chains_total chains on its slice of the observations."""
    if shard == "chains":
        off, cnt = shard_chains(n_chains_total, rank, world)
        return CUDAMCMCBackend(n_chains=cnt, chain_offset=off, device=local_rank, rank=rank,
                               world_size=world, shard_mode="chains", **kw)
    return CUDAMCMCBackend(n_chains=n_chains_total, chain_offset=0, device=local_rank, rank=rank,
                           world_size=world, shard_mode="obs", comm_id=comm_id, **kw)


def gather_chain_axis(dist, array, dst=0):
    """Concatenate per-rank arrays along their last (chain) axis on rank `dst`."""
    parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(np.ascontiguousarray(array), parts, dst=dst)
    return np.concatenate(parts, axis=-1) if parts is not None else None
