"""Generates tests/golden/oracle_cfg_small.json from the CPU oracle (the reference itself
cannot run here: Julia is not installed).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import extensiblemcmc_jl_b200 as em  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.parity import cfg2_updates, theta_init_for  # noqa: E402

rng = np.random.default_rng(2)
x = 1.5 + 2.0 * rng.standard_normal(301)          # odd N on purpose
kw = dict(eps0=0.2, scale=0.02, k=5, vmin=1e-7, offset=2.0)
Cn, M, seed = 6, 30, 12345
th0 = theta_init_for(x, Cn)
o = orc.Oracle(em.GsnTargetLaw([0.0]), cfg2_updates(**kw), x, th0, n_chains=Cn, seed=seed)
r = o.run(list(em.MCMCSchedule(M, 2)))
out = dict(obs=x.tolist(), update_kwargs=kw, n_chains=Cn, n_iters=M, seed=seed, theta_init=th0.tolist(),
           accepted=r["accepted"].tolist(), theta=r["theta"].tolist(), theta_prop=r["theta_prop"].tolist(),
           ll=r["ll"].tolist(), ll_prop=r["ll_prop"].tolist(), llr=r["llr"].tolist(),
           proposals=r["proposals"].tolist(), exp_draws=r["exp_draws"].tolist(),
           eps1=o.eps(1).tolist(), eps2=o.eps(2).tolist())
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_cfg_small.json"), "w"))
print("written; accept rate", r["accepted"].mean())
