"""run! and the block driver that replaces __run! (src/run.jl:34-83).

The reference executes one (mcmciter, pidx) schedule element at a time on the host.  Here
the host only walks the schedule: consecutive elements are collected into blocks and each
block is ONE call into libextmcmc_cuda (`extmcmc_run_block`, one CUDA graph launch)
which performs, for every chain, update_workspaces! -> update! -> update_adaptation!
(src/run.jl:101-208).  A block ends where a callback's `check_if_execute` fires (the
reference queries it before and after every step, src/callbacks.jl:33-45), so callbacks
and `reschedule!` observe exactly the state they would observe in the reference.
"""
import ctypes as C

import numpy as np

from . import _abi
from .mcmc import init_
from .schedule import Step
from .types import POSTSTEP, PRESTEP
from . import workspaces as W


def run_(mcmc, num_mcmc_steps, data, theta_init, callbacks=(), **kwargs):
    """run!(mcmc, num_mcmc_steps, data, theta_init, callbacks; kwargs...) -- run.jl:34-54."""
    callbacks = list(callbacks)
    init_(mcmc, num_mcmc_steps, data, theta_init, kwargs.get("exclude_updates", ()), **kwargs)
    local_wss = W.create_workspaces(mcmc.backend, mcmc)
    for cb in callbacks:
        cb.init_(mcmc.workspace)
    __run_(mcmc.workspace, local_wss, mcmc.updates, mcmc.schedule, callbacks)
    final = Step(None, None, num_mcmc_steps, None)
    for cb in callbacks:
        cb.cleanup_(mcmc.workspace, local_wss, final)
    return mcmc.workspace, local_wss


def _flush(ws, local_wss, block):
    """Ship one block of schedule elements to the device and mirror its history rows."""
    if not block:
        return
    n = len(block)
    arr = (_abi.Step * n)()
    for k, s in enumerate(block):
        arr[k].mcmciter = s.mcmciter
        arr[k].prev_mcmciter = s.prev_mcmciter if s.prev_mcmciter is not None else 0
        arr[k].pidx = s.pidx - 1
        arr[k].prev_pidx = (s.prev_pidx - 1) if s.prev_pidx is not None else -1
    ws._ck(ws.lib.extmcmc_run_block(ws.handle, arr, n))
    ws.pending.append((ws.seq_launched, list(block)))
    ws.seq_launched += n


def _drain(ws, local_wss):
    """Copy the history rows of all launched blocks back (blocks until they finished)."""
    lib = ws.lib
    for seq_lo, block in ws.pending:
        n = len(block)
        if ws.keep_history:
            p, Cn = ws.p, ws.C
            th = np.empty((n, p, Cn)); thp = np.empty((n, p, Cn))
            l = np.empty((n, Cn)); lp = np.empty((n, Cn))
            acc = np.empty((n, Cn), dtype=np.uint8)
            ws._ck(lib.extmcmc_get_history(ws.handle, seq_lo, seq_lo + n, _abi.dptr(th), _abi.dptr(thp),
                                           _abi.dptr(l), _abi.dptr(lp), acc.ctypes.data_as(_abi.c_uint8_p)))
            it = np.fromiter((s.mcmciter - 1 for s in block), dtype=np.int64, count=n)
            pj = np.fromiter((s.pidx - 1 for s in block), dtype=np.int64, count=n)
            ws.sub_ws.state_history[it, pj] = th
            ws.sub_ws.state_proposal_history[it, pj] = thp
            for j in np.unique(pj):
                m = pj == j
                lw = local_wss[j]
                lw.sub_ws.ll_history[it[m], 0] = l[m]
                lw.sub_ws_prop.ll_history[it[m], 0] = lp[m]
                lw.acceptance_history[it[m]] = acc[m].astype(bool)
            last = {}
            for k, s in enumerate(block):
                last[s.pidx - 1] = k
            for j, k in last.items():
                lw = local_wss[j]
                idx = np.asarray(ws.updates[j].coords) - 1
                lw.sub_ws.state[:] = th[k][idx]
                lw.sub_ws_prop.state[:] = thp[k][idx]
                lw.sub_ws.ll[0] = l[k]
                lw.sub_ws_prop.ll[0] = lp[k]
            ws.sub_ws.state = th[-1].copy()
    ws.pending.clear()
    ws.sync()              # raises on a domain error (the reference would have thrown)
    if not ws.keep_history:
        ws.refresh_state()


def __run_(global_ws, local_wss, updates, schedule, callbacks):
    """__run!(global_ws, local_wss, updates, schedule, callbacks) -- run.jl:64-83."""
    ws = global_ws
    ws.updates = updates
    ws.pending = []
    ws.seq_launched = 0
    block = []
    for step in schedule:
        pre = [cb for cb in callbacks if cb.check_if_execute(step, PRESTEP)]
        if pre:
            _flush(ws, local_wss, block); block = []
            _drain(ws, local_wss)
            for cb in pre:
                cb.execute_(ws, local_wss, step, PRESTEP)
        block.append(step)
        post = [cb for cb in callbacks if cb.check_if_execute(step, POSTSTEP)]
        if post:
            _flush(ws, local_wss, block); block = []
            _drain(ws, local_wss)
            for cb in post:
                cb.execute_(ws, local_wss, step, POSTSTEP)
        elif len(block) >= ws.block_len:
            _flush(ws, local_wss, block); block = []
            if len(ws.pending) >= 2:          # ring holds two blocks: drain the older one(s)
                _drain(ws, local_wss)
    _flush(ws, local_wss, block)
    _drain(ws, local_wss)
