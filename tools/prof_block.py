"""Small driver for profiling the persistent block kernels under ncu (one workload, a few blocks).
usage: python tools/prof_block.py cfg2|cfg4|cfg5 [iterations] [block_iters]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import extensiblemcmc_jl_b200 as em  # noqa: E402
from extensiblemcmc_jl_b200 import _abi  # noqa: E402
from extensiblemcmc_jl_b200.mcmc import init_  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sv = int(os.environ.get("SWEEP_VARIANT", "0"))
if what == "cfg2":
    x = bench.cfg2_data()
    ws = bench.make_ws(em, x, 4096, 0, 0, block_len=2 * blk, use_graphs=True, sweep_variant=sv)
    nu = 2
elif what == "cfg5":
    n = int(os.environ.get("N_OBS", str(1 << 28)))
    m = em.MCMC(bench.cfg2_updates(em), backend=em.CUDAMCMCBackend(n_chains=8, seed=6, history="none", block_len=2 * blk,
                                                                   use_graphs=True, sweep_variant=sv))
    init_(m, 1, dict(P=em.GsnTargetLaw([0.0]), obs=em.DeviceGeneratedObs(n, 1.5, 2.0, 6)), np.repeat(np.array([[1.5], [4.0]]), 8, axis=1))
    ws, nu = m.workspace, 2
else:
    rng = np.random.default_rng(5)
    C, G, ng = 8192, 8, 4096
    tg = rng.standard_normal(G)
    yv = np.concatenate([tg[g] + rng.standard_normal(ng) for g in range(G)])
    ups = [em.MALAUpdate(0.02, list(range(1, G + 1)), adpt=em.AdaptationMALA(adapt_every_k_steps=50, scale=1e-3, min=1e-5)),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.3]), [G + 1], adpt=em.AdaptationUnifRW([0.0], adapt_every_k_steps=50, scale=0.02)),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.3], [True]), [G + 2], prior=em.ImproperPosPrior(),
                               adpt=em.AdaptationUnifRW([0.0], adapt_every_k_steps=50, scale=0.02))]
    m = em.MCMC(ups, backend=em.CUDAMCMCBackend(n_chains=C, seed=7, history="none", block_len=3 * blk, use_graphs=True, sweep_variant=sv))
    init_(m, 1, dict(P=em.HierNormalLaw(G), obs=yv, groups=np.repeat(np.arange(G), ng)), np.concatenate([np.zeros(G), [0.0, 1.0]]))
    ws, nu = m.workspace, 3
print("kernel:", ws.lib.extmcmc_sweep_variant_name(ws.handle).decode())
it = 1
import ctypes


def block(it0):
    arr = (_abi.Step * (nu * blk))()
    k = 0
    for it in range(it0, it0 + blk):
        for pj in range(nu):
            first = it == 1 and pj == 0
            arr[k].mcmciter, arr[k].pidx = it, pj
            arr[k].prev_pidx = -1 if first else (pj - 1 if pj else nu - 1)
            arr[k].prev_mcmciter = 0 if first else (it if pj else it - 1)
            k += 1
    return arr


for _ in range(2):
    ws._ck(ws.lib.extmcmc_run_block(ws.handle, block(it), nu * blk)); it += blk
ws.sync()
ws._ck(ws.lib.extmcmc_event_record(ws.handle, 0))
for _ in range(iters):
    ws._ck(ws.lib.extmcmc_run_block(ws.handle, block(it), nu * blk)); it += blk
ws._ck(ws.lib.extmcmc_event_record(ws.handle, 1))
ws.sync()
ms = ctypes.c_float()
ws._ck(ws.lib.extmcmc_event_elapsed(ws.handle, 0, 1, ctypes.byref(ms)))
print(f"{what}: {ms.value / (iters * blk):.4f} ms per iteration ({iters} blocks of {blk} iterations)")
ws.close()
