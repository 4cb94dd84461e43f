"""The library's launch-time variants must not change a single bit of the results: the lean
step-kernel instantiations (EXTMCMC_LEAN, csrc/step_device.cuh SpecLean) against the general ones,
the L2 evict_first hint on the observation stream (EXTMCMC_L2_HINT) and the programmatic-dependent-
launch masks (EXTMCMC_PDL).  The switches are read once per process, so every variant runs in a
process of its own; each prints a digest of the final state, step sizes, moments and counters of two
jobs (cfg 2 shape: two uniform walks; cfg 4 shape: MALA + two walks on the hierarchical law) under the
device's own Philox stream."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_JOB = r"""
import hashlib, json, sys
import numpy as np
sys.path.insert(0, %r)
import extensiblemcmc_jl_b200 as em
from tests.parity import GpuSession

def digest(s, hist, nu):
    st = s.stats()
    h = hashlib.sha256()
    parts = [hist["theta"], hist["theta_prop"], hist["ll"], hist["ll_prop"], hist["accepted"], st["mean"], st["cov"],
             st["rolling_ar"], st["n_accept"], st["n_prop"]] + [s.eps(u) for u in range(1, nu + 1)]
    for a in parts:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()

out = {}
rng = np.random.default_rng(1)
# cfg 2 shape, few and many chains (obs mapping with the fused tail / chains mapping), graphs on
for C, N in ((8, 300000), (1024, 20000)):
    x = 1.5 + 2.0 * rng.standard_normal(N)
    mk = lambda: em.AdaptationUnifRW([0.0], adapt_every_k_steps=10, scale=5e-3, min=1e-7, max=1e7, offset=1.0)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.05]), [1], adpt=mk()),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.05], [True]), [2], prior=em.ImproperPosPrior(), adpt=mk())]
    s = GpuSession(em.GsnTargetLaw([0.0]), ups, x, np.array([1.4, 4.2]), C, seed=5, n_steps_hint=120, use_graphs=1)
    out["cfg2_C%%d" %% C] = digest(s, s.run(list(em.MCMCSchedule(60, 2))), 2)
    s.close()
# cfg 4 shape
G, ng, C = 4, 500, 512
tg = rng.standard_normal(G)
yv = np.concatenate([tg[g] + rng.standard_normal(ng) for g in range(G)])
ups = [em.MALAUpdate(0.05, list(range(1, G + 1)), adpt=em.AdaptationMALA(adapt_every_k_steps=10, scale=1e-3, min=1e-5)),
       em.RandomWalkUpdate(em.UniformRandomWalk([0.3]), [G + 1], adpt=em.AdaptationUnifRW([0.0], adapt_every_k_steps=10, scale=0.02)),
       em.RandomWalkUpdate(em.UniformRandomWalk([0.3], [True]), [G + 2], prior=em.ImproperPosPrior(),
                           adpt=em.AdaptationUnifRW([0.0], adapt_every_k_steps=10, scale=0.02))]
s = GpuSession(em.HierNormalLaw(G), ups, yv, np.concatenate([np.zeros(G), [0.0, 1.0]]), C, seed=9, n_steps_hint=120,
               use_graphs=1, y=np.repeat(np.arange(G), ng).astype(np.float64))
out["cfg4"] = digest(s, s.run(list(em.MCMCSchedule(40, 3))), 3)
s.close()
print(json.dumps(out))
""" % ROOT


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", _JOB], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])


@pytest.mark.gpu
def test_launch_variants_are_bit_identical():
    base = _run({})
    assert len(base) == 3
    for env in ({"EXTMCMC_LEAN": "0"}, {"EXTMCMC_DEFER": "0"}, {"EXTMCMC_L2_HINT": "0"}, {"EXTMCMC_PDL": "0"}, {"EXTMCMC_PDL": "7"}):
        assert _run(env) == base, env
