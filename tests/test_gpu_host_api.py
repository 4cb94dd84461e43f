"""The reference-facing host API (MCMC / run! / callbacks / workspaces) on the GPU path."""
import ctypes as C
import io
import os

import numpy as np
import pytest

import extensiblemcmc_jl_b200 as em
from extensiblemcmc_jl_b200 import _abi
from tests.parity import GpuSession

pytestmark = pytest.mark.gpu


def _setup(n_chains, M, **bk):
    rng = np.random.default_rng(10)
    x = 1.0 + 2.0 * rng.standard_normal(500)
    mk = lambda: em.AdaptationUnifRW([0.0], adapt_every_k_steps=20, scale=0.05)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.5]), [1], adpt=mk()),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.5], [True]), [2], prior=em.ImproperPosPrior(), adpt=mk())]
    mcmc = em.MCMC(ups, backend=em.CUDAMCMCBackend(n_chains=n_chains, seed=4, **bk))
    return mcmc, ups, x


def test_reference_smoke_shape():
    # the reference's own "mcmc" testset shape (test/runtests.jl:87-114), d = 1 variant:
    # 2 single-site uniform walks, 1000 iterations, one chain, no assertion on the values
    mcmc, ups, x = _setup(1, 1000)
    ws, lws = em.run_(mcmc, 1000, dict(P=em.GsnTargetLaw([0.0]), obs=x[:10]), [0.0, 1.0], [])
    assert ws.sub_ws.state_history.shape == (1000, 2, 2, 1)
    assert not np.isnan(ws.sub_ws.state_history).any()
    assert len(lws) == 2 and lws[0].acceptance_history.shape == (1000, 1)
    assert lws[0].acceptance_history[0, 0]                       # first proposal always accepted
    # the accessors of src/workspaces.jl:91-136,294-385 on the finished run
    assert em.num_mcmc_steps(ws) == 1000 and em.num_updt(ws) == 2
    assert em.name_of_update(lws[0]) == "RandomWalkUpdate"
    assert em.state(ws).shape == (2, 1) and em.state(lws[1]).shape == (1, 1)
    assert em.estim_mean(ws).shape == (2, 1) and em.estim_cov(ws).shape == (2, 2, 1)
    assert bool(em.accepted(lws[0], 1)[0]) is True
    assert np.isfinite(em.ll(lws[0], 500)).all() and np.isfinite(em.ll_prop(lws[0], 500)).all()
    assert em.llr(lws[0], 500).shape == (1,)
    em.set_accepted_(lws[0], 1, False)                            # set_accepted! :299-303
    assert not lws[0].acceptance_history[0, 0]
    ws.close()


def test_run_matches_raw_abi_and_block_length_is_irrelevant():
    outs = []
    for bl in (7, 64):
        mcmc, ups, x = _setup(32, 50, block_len=bl)
        ws, lws = em.run_(mcmc, 50, dict(P=em.GsnTargetLaw([0.0]), obs=x), [0.0, 1.0])
        outs.append((ws.sub_ws.state_history.copy(), lws[1].sub_ws.ll_history.copy(),
                     lws[0].acceptance_history.copy(), ws.stats()))
        ws.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][3]["mean"], outs[1][3]["mean"])
    s = GpuSession(em.GsnTargetLaw([0.0]), ups, x, [0.0, 1.0], 32, seed=4, n_steps_hint=100)
    r = s.run(list(em.MCMCSchedule(50, 2)))
    assert np.array_equal(r["theta"].reshape(50, 2, 2, 32), outs[0][0])
    assert np.array_equal(r["accepted"].reshape(50, 2, 32)[:, 0].astype(bool), outs[0][2])
    s.close()


def test_callbacks_and_csv(tmp_path):
    mcmc, ups, x = _setup(3, 40, block_len=16)
    buf = io.StringIO()
    cbs = [em.SavingCallback(save_at_iters=[10, 25], filename="chain", path=str(tmp_path)),
           em.REPLCallback(print_every_k_iter=20, file=buf)]
    seen = []

    class Probe(em.Callback):
        def check_if_execute(self, step, flag):
            return isinstance(flag, em.PostMCMCStep) and step.mcmciter == 13 and step.pidx == 1
        def execute_(self, ws, lws, step, flag):
            # at this point every earlier step must be mirrored on the host
            e = ws.executed
            seen.append((bool(e[:12].all()), bool(e[12, 0]), bool(not e[12, 1:].any() and not e[13:].any())))
    ws, lws = em.run_(mcmc, 40, dict(P=em.GsnTargetLaw([0.0]), obs=x), [0.0, 1.0], cbs + [Probe()])
    assert seen == [(True, True, True)]
    out = buf.getvalue()
    assert "Initializing an MCMC chain" in out and "20.1 RandomWalkUpdate" in out and "40.2" in out
    files = sorted(os.listdir(tmp_path))
    assert files == ["chain_chain0.csv", "chain_chain1.csv", "chain_chain2.csv"]
    rows = open(tmp_path / "chain_chain1.csv").read().strip().split("\n")
    # iterations 1..9 (save at 10), 10..24 (save at 25), then the end-of-run save from the last
    # save point to M - 1 (callbacks.jl:223,234-239)
    its = [int(r.split(",")[0]) for r in rows]
    assert its[0] == 1 and its[-1] == 39 and len(rows) == 2 * 39
    f = rows[0].split("!")
    assert f[0].strip() == "1, 1," and len(f) == 6
    th = [float(v) for v in f[1].split(",") if v.strip()]
    assert th == list(ws.sub_ws.state_history[0, 0, :, 1])
    assert f[5].strip() in (",true,", ",false,")
    ws.close()


def test_reschedule_from_a_callback():
    mcmc, ups, x = _setup(8, 12, block_len=5)

    class Drop(em.Callback):
        def check_if_execute(self, step, flag):
            return isinstance(flag, em.PostMCMCStep) and step.mcmciter == 4 and step.pidx == 1
        def execute_(self, ws, lws, step, flag):
            em.reschedule_(mcmc.schedule, 0, [2])
    ws, lws = em.run_(mcmc, 12, dict(P=em.GsnTargetLaw([0.0]), obs=x), [0.0, 1.0], [Drop()])
    h = ws.sub_ws.state_history
    # update 2 still runs at (4, 2) (the next state was already computed, schedule.jl:56-66)
    assert not np.isnan(h[3, 1]).any() and np.isnan(h[4:, 1]).all() and not np.isnan(h[:, 0]).any()
    assert (ws.stats()["n_prop"][:, 0] == [12, 4]).all()
    ws.close()


def test_checkpoint_resume_continues_the_same_chains_bit_for_bit():
    """extmcmc_checkpoint_save / _load + MCMCSchedule(...; start = ...) (src/schedule.jl:28): a run cut in
    two, with the second half in a FRESH handle, equals the uninterrupted run -- state, step sizes,
    running moments, rolling acceptance rates and decision counts, bit for bit."""
    from tests.parity import GpuSession, cfg2_updates, theta_init_for
    x = 1.5 + 2.0 * np.random.default_rng(5).standard_normal(3000)
    Cn, M = 70, 60
    th0 = theta_init_for(x, Cn)
    mk = lambda: cfg2_updates(eps0=0.05, scale=5e-3, k=7, offset=2.0)
    steps = list(em.MCMCSchedule(M, 2, [(2, range(20, 35))]))
    full = GpuSession(em.GsnTargetLaw([0.0]), mk(), x, th0, Cn, seed=21, n_steps_hint=len(steps), roll_window=10)
    rf = full.run(steps)
    cut = 47                                                    # in the middle of an iteration
    a = GpuSession(em.GsnTargetLaw([0.0]), mk(), x, th0, Cn, seed=21, n_steps_hint=len(steps), roll_window=10)
    ra = a.run(steps[:cut])
    n = C.c_int64()
    a.ck(a.lib.extmcmc_checkpoint_size(a.h, C.byref(n)))
    blob = (C.c_uint8 * n.value)()
    a.ck(a.lib.extmcmc_checkpoint_save(a.h, blob, n.value))
    a.close()
    b = GpuSession(em.GsnTargetLaw([0.0]), mk(), x, th0 * 0 + 1.0, Cn, seed=21, n_steps_hint=len(steps), roll_window=10)
    b.ck(b.lib.extmcmc_checkpoint_load(b.h, blob, n.value))
    b.seq = cut
    rb = b.run(steps[cut:])
    for k in ("theta", "theta_prop", "ll", "accepted"):
        assert np.array_equal(rf[k], np.concatenate([ra[k], rb[k]])), k
    sf, sb = full.stats(), b.stats()
    for k in ("mean", "cov", "rolling_ar", "n_accept", "n_prop"):
        assert np.array_equal(sf[k], sb[k]), k
    assert np.array_equal(full.eps(1), b.eps(1)) and np.array_equal(full.eps(2), b.eps(2))
    # a blob of another shape is refused
    c = GpuSession(em.GsnTargetLaw([0.0]), mk(), x, th0[:, :8], 8, seed=21)
    assert c.lib.extmcmc_checkpoint_load(c.h, blob, n.value) == _abi.EINVAL
    for s in (full, b, c):
        s.close()
