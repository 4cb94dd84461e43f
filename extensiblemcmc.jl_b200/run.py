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
    """Ship one block of schedule elements to the device (asynchronous) and start the
    copy-back of its history rows; while the GPU runs it, the previous block's rows are
    moved from the pinned staging area into the host histories."""
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
    seq_lo = ws.seq_launched
    ws.seq_launched += n
    if ws.keep_history:
        _finish_fetch(ws, local_wss)                       # rows of the previous block
        ws._ck(ws.lib.extmcmc_history_fetch_begin(ws.handle, seq_lo, seq_lo + n))
        ws.fetching = list(block)


def _finish_fetch(ws, local_wss):
    block = ws.fetching
    if block is None:
        return
    ws.fetching = None
    lib, n, p, Cn, NU = ws.lib, len(block), ws.p, ws.C, ws.NU
    it = np.fromiter((s.mcmciter - 1 for s in block), dtype=np.int64, count=n)
    pj = np.fromiter((s.pidx - 1 for s in block), dtype=np.int64, count=n)
    flat = it * NU + pj
    sh, sph = ws.sub_ws.state_history, ws.sub_ws.state_proposal_history
    if np.array_equal(flat, flat[0] + np.arange(n)):
        # dense block: the rows are consecutive in the [M][NU] host arrays -- copy straight in
        f0 = int(flat[0])
        sl = lambda a: a.reshape((-1,) + a.shape[2:])[f0:f0 + n]
        th, thp, l, lp, acc = sl(sh), sl(sph), sl(ws.ll_all), sl(ws.llp_all), sl(ws.acc_all)
        ws._ck(lib.extmcmc_history_fetch_end(ws.handle, _abi.dptr(th), _abi.dptr(thp), _abi.dptr(l),
                                             _abi.dptr(lp), acc.ctypes.data_as(_abi.c_uint8_p)))
    else:
        th = np.empty((n, p, Cn)); thp = np.empty((n, p, Cn))
        l = np.empty((n, Cn)); lp = np.empty((n, Cn))
        acc = np.empty((n, Cn), dtype=np.uint8)
        ws._ck(lib.extmcmc_history_fetch_end(ws.handle, _abi.dptr(th), _abi.dptr(thp), _abi.dptr(l),
                                             _abi.dptr(lp), acc.ctypes.data_as(_abi.c_uint8_p)))
        sh[it, pj] = th
        sph[it, pj] = thp
        ws.ll_all[it, pj] = l
        ws.llp_all[it, pj] = lp
        ws.acc_all[it, pj] = acc
    ws.executed[it, pj] = True
    # "current" values of every local workspace = its last executed step in this block
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
    ws.sub_ws.state = np.array(th[n - 1])


def _drain(ws, local_wss):
    """Host-visible point: every launched step has finished and is mirrored on the host."""
    if ws.keep_history:
        _finish_fetch(ws, local_wss)
    ws.sync()              # raises on a domain error (the reference would have thrown)
    if not ws.keep_history:
        ws.refresh_state()


def __run_(global_ws, local_wss, updates, schedule, callbacks):
    """__run!(global_ws, local_wss, updates, schedule, callbacks) -- run.jl:64-83."""
    ws = global_ws
    ws.updates = updates
    ws.fetching = None
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
    _flush(ws, local_wss, block)
    _drain(ws, local_wss)
    if ws.keep_history and not ws.executed.all():
        never = ~ws.executed
        ws.sub_ws.state_history[never] = np.nan
        ws.sub_ws.state_proposal_history[never] = np.nan
