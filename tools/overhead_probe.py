"""Fixed per-update-step overhead of the block driver: cfg-2 chains with a tiny dataset, so the
sweep is negligible and what remains is proposal + accept + launch gaps (graph mode)."""
import ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import extensiblemcmc_jl_b200 as em
from extensiblemcmc_jl_b200 import _abi
import bench

for C, N in [(4096, 2048), (8192, 32768), (65536, 2048)]:
    x = 1.5 + 2.0 * np.random.default_rng(2).standard_normal(N)
    bench.N_OBS = N
    for blk in (1, 64):
        from extensiblemcmc_jl_b200.mcmc import init_
        mcmc = em.MCMC(bench.cfg2_updates(em), backend=em.CUDAMCMCBackend(n_chains=C, seed=3, history="none",
                                                                         block_len=2 * blk, use_graphs=True))
        init_(mcmc, 1, dict(P=em.GsnTargetLaw([0.0]), obs=x), bench.cfg2_theta_init(x, C))
        ws = mcmc.workspace
        lib, h = ws.lib, ws.handle
        it = 1
        for _ in range(3):
            ws._ck(lib.extmcmc_run_block(h, bench.steps_for(None, _abi, it, blk), 2 * blk)); it += blk
        ws.sync()
        reps = 50
        ws._ck(lib.extmcmc_event_record(h, 0))
        for _ in range(reps):
            ws._ck(lib.extmcmc_run_block(h, bench.steps_for(None, _abi, it, blk), 2 * blk)); it += blk
        ws._ck(lib.extmcmc_event_record(h, 1))
        ms = ctypes.c_float()
        ws._ck(lib.extmcmc_event_elapsed(h, 0, 1, ctypes.byref(ms)))
        print(json.dumps(dict(C=C, N=N, iters_per_graph=blk, us_per_update_step=ms.value * 1e3 / (reps * blk * 2),
                              variant=lib.extmcmc_sweep_variant_name(h).decode())))
        ws.close()
