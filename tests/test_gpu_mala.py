"""MALAUpdate (a #TODO stub in the reference, src/updates.jl:216-218; build-defined semantics),
device gradients and the hierarchical-normal law of BASELINE cfg 4, against the oracle.
No reference oracle exists for these: pinned by finite differences, by the
acceptance-rate -> 1 limit as tau -> 0 (detailed balance of the proposal/Hastings pair) and by
posterior agreement with an independent long oracle run."""
import math

import numpy as np
import pytest

import extensiblemcmc_jl_b200 as em
from extensiblemcmc_jl_b200 import _abi
from oracle import oracle as orc
from tests.parity import GpuSession, replay_compare, theta_init_for

pytestmark = pytest.mark.gpu


def _hier_data(G=8, ng=400, seed=5, ragged=False):
    rng = np.random.default_rng(seed)
    tg = rng.standard_normal(G)
    sizes = [ng + (3 * g + 1 if ragged else 0) for g in range(G)]
    y = np.concatenate([tg[g] + rng.standard_normal(sizes[g]) for g in range(G)])
    grp = np.concatenate([np.full(sizes[g], g) for g in range(G)])
    return y, grp, tg


def _hier_updates(G, tau=0.1):
    return [em.MALAUpdate(tau, list(range(1, G + 1)),
                          adpt=em.AdaptationMALA(adapt_every_k_steps=6, scale=0.004, offset=1.0)),
            em.RandomWalkUpdate(em.UniformRandomWalk([0.4]), [G + 1],
                                adpt=em.AdaptationUnifRW([0.0], adapt_every_k_steps=5, scale=0.02, offset=1.0)),
            em.RandomWalkUpdate(em.UniformRandomWalk([0.4], [True]), [G + 2], prior=em.ImproperPosPrior())]


def _hier_theta0(G, n_chains, seed=2):
    rng = np.random.default_rng(seed)
    th = np.zeros((G + 2, n_chains))
    th[:G] = 0.3 * rng.standard_normal((G, n_chains))
    th[G] = 0.1 * rng.standard_normal(n_chains)
    th[G + 1] = np.exp(0.2 * rng.standard_normal(n_chains))
    return th


def _clean(rep):
    assert rep["accept_mismatch"] == 0 and rep["near_ties"] == 0, rep
    assert rep["theta_bitexact"] and rep["ll_rel_err"] < 1e-10, rep
    assert rep["eps_bitexact"] and rep["mean_bitexact"] and rep["cov_bitexact"] and rep["counts_equal"], rep
    assert rep["rolling_ar_bitexact"] and rep["final_state_bitexact"], rep
    assert 0.02 < rep["accept_rate"] < 0.98, rep


@pytest.mark.parametrize("n_chains", [70, 5])
def test_gradient_gsn1d_against_oracle_and_finite_differences(n_chains):
    x = 1.5 + 2.0 * np.random.default_rng(1).standard_normal(3001)
    law = em.GsnTargetLaw([0.0])
    th0 = theta_init_for(x, n_chains)
    ups = [em.MALAUpdate(0.01, [1])]
    g = GpuSession(law, ups, x, th0, n_chains)
    ll, gr = g.eval_grad()
    o = orc.Oracle(law, ups, x, th0, n_chains)
    llo, gro = o.loglik_grad(th0)
    assert np.allclose(ll, llo, rtol=1e-10, atol=0) and np.allclose(gr, gro, rtol=1e-9, atol=1e-9)
    h = 1e-6
    for k in range(2):
        e = np.zeros_like(th0); e[k] = h
        fd = (o.loglik(th0 + e) - o.loglik(th0 - e)) / (2 * h)
        assert np.allclose(gr[k], fd, rtol=1e-5, atol=1e-4)
    g.close()


@pytest.mark.parametrize("n_chains,ragged", [(300, False), (6, True)])
def test_gradient_hier_against_oracle_and_finite_differences(n_chains, ragged):
    G = 8
    y, grp, _ = _hier_data(G, 333, ragged=ragged)
    law = em.HierNormalLaw(G)
    th0 = _hier_theta0(G, n_chains)
    ups = _hier_updates(G)
    g = GpuSession(law, ups, y, th0, n_chains, y=grp)
    ll, gr = g.eval_grad()
    o = orc.Oracle(law, ups, y, th0, n_chains, y=grp)
    llo, gro = o.loglik_grad(th0)
    assert np.allclose(ll, llo, rtol=1e-10, atol=0)
    assert np.allclose(gr, gro, rtol=1e-9, atol=1e-9)
    assert np.allclose(g.eval_loglik(), llo, rtol=1e-10, atol=0)
    h = 1e-6
    for k in (0, G - 1, G, G + 1):
        e = np.zeros_like(th0); e[k] = h
        fd = (o.loglik(th0 + e) - o.loglik(th0 - e)) / (2 * h)
        assert np.allclose(gr[k], fd, rtol=1e-5, atol=1e-4)
    g.close()


def test_mala_replay_parity_gsn1d():
    x = 1.5 + 2.0 * np.random.default_rng(3).standard_normal(2500)
    ups = [em.MALAUpdate(0.05, [1], prior=em.StandardPrior(em.Normal(0.0, 10.0)),
                         adpt=em.AdaptationMALA(adapt_every_k_steps=5, scale=0.003, offset=1.0)),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.05], [True]), [2], prior=em.ImproperPosPrior())]
    rep = replay_compare(x, 90, 50, seed=6, updates=ups, block=17, history_window=32)
    _clean(rep)


def test_mala_replay_parity_joint_update():
    x = 1.5 + 2.0 * np.random.default_rng(4).standard_normal(1800)
    ups = [em.MALAUpdate(0.04, [1, 2], adpt=em.AdaptationMALA(adapt_every_k_steps=7, scale=0.002, offset=1.0))]
    rep = replay_compare(x, 64, 60, seed=9, updates=ups)
    _clean(rep)


@pytest.mark.parametrize("n_chains,ragged", [(200, False), (7, True)])
def test_cfg4_schedule_replay_parity(n_chains, ragged):
    # BASELINE cfg 4: MALA on theta_1..8, uniform walk on mu, multiplicative walk on tau
    G = 8
    y, grp, _ = _hier_data(G, 256, ragged=ragged)
    rep = replay_compare(y, n_chains, 30, seed=12, updates=_hier_updates(G), law=em.HierNormalLaw(G), y=grp,
                         theta_init=_hier_theta0(G, n_chains), exclude=[(2, range(5, 9))], block=31,
                         history_window=40)
    _clean(rep)


def test_cfg4_schedule_replay_parity_variances_only():
    # stats_mode = 1 (running mean + variances; what bench.py's cfg3 run uses): the variances are
    # the diagonal of the reference's covariance recursion, bit for bit
    G = 8
    y, grp, _ = _hier_data(G, 256)
    rep = replay_compare(y, 150, 20, seed=13, updates=_hier_updates(G), law=em.HierNormalLaw(G), y=grp,
                         theta_init=_hier_theta0(G, 150), history_window=64, stats_mode=1)
    _clean(rep)


def test_mala_accepts_everything_as_tau_goes_to_zero():
    G = 4
    y, grp, _ = _hier_data(G, 100, seed=8)
    law = em.HierNormalLaw(G)
    th0 = _hier_theta0(G, 64)
    rates = []
    for tau in (0.3, 1e-3):
        g = GpuSession(law, [em.MALAUpdate(tau, [1, 2, 3, 4])], y, th0, 64, seed=4, y=grp, n_steps_hint=40)
        r = g.run(list(em.MCMCSchedule(40, 1)))
        rates.append(r["accepted"][1:].mean())
        g.close()
    assert rates[1] > 0.995 and rates[0] < rates[1]


def test_cfg4_posterior_matches_oracle_and_graphs_and_sharding_are_exact():
    G = 8
    y, grp, _ = _hier_data(G, 400, seed=11)
    law = em.HierNormalLaw(G)
    data = dict(P=law, obs=y, groups=grp)
    th0 = np.concatenate([np.zeros(G), [0.0, 1.0]])
    mk = lambda: [em.MALAUpdate(0.1, list(range(1, G + 1)),
                                adpt=em.AdaptationMALA(adapt_every_k_steps=25, scale=0.01, offset=2.0)),
                  em.RandomWalkUpdate(em.UniformRandomWalk([0.5]), [G + 1]),
                  em.RandomWalkUpdate(em.UniformRandomWalk([0.5], [True]), [G + 2], prior=em.ImproperPosPrior())]
    Cn, M = 256, 1500
    mcmc = em.MCMC(mk(), backend=em.CUDAMCMCBackend(n_chains=Cn, seed=31, block_len=96))
    ws, lws = em.run_(mcmc, M, data, th0)
    tr = ws.sub_ws.state_history[500:, 2]                     # [iters, p, C]
    o = orc.Oracle(law, mk(), y, th0, 32, seed=77, y=grp)
    ro = o.run(list(em.MCMCSchedule(M, 3)), n_threads=8, record=False)
    tro = ro["theta"].reshape(M, 3, G + 2, 32)[500:, 2]
    ess, esso = em.ess_geyer(tr), em.ess_geyer(tro)
    for k in range(G + 2):
        se = math.sqrt(tr[:, k].var() / ess[k].sum() + tro[:, k].var() / esso[k].sum())
        assert abs(tr[:, k].mean() - tro[:, k].mean()) < 3 * se, k
    acc = ws.stats()["n_accept"].sum(axis=1) / ws.stats()["n_prop"].sum(axis=1)
    assert 0.35 < acc[0] < 0.8                                 # MALA step adapted towards 0.574
    full = ws.sub_ws.state_history[:60].copy()
    tau = ws.eps(1).copy()
    ws.close()
    # eager launches and a 2-way chain split reproduce the graph run bit-for-bit
    outs = []
    for off, cnt, graphs in ((0, Cn, False), (0, 100, True), (100, Cn - 100, True)):
        m2 = em.MCMC(mk(), backend=em.CUDAMCMCBackend(n_chains=cnt, chain_offset=off, seed=31, block_len=50,
                                                      use_graphs=graphs))
        w2, _ = em.run_(m2, 60, data, th0)
        outs.append(w2.sub_ws.state_history.copy())
        w2.close()
    assert np.array_equal(outs[0], full)
    assert np.array_equal(np.concatenate(outs[1:], axis=-1), full)
    assert tau.shape == (1, Cn)


def test_mala_unsupported_combinations():
    x = np.zeros(10)
    with pytest.raises(_abi.ExtMCMCError) as ei:       # no device gradient for the general-d law
        GpuSession(em.GsnTargetLaw(np.zeros(2)), [em.MALAUpdate(0.1, [1])], np.zeros((10, 2)),
                   em.GsnTargetLaw(np.zeros(2)).theta, 4)
    assert ei.value.code == _abi.EUNSUPPORTED
    with pytest.raises(_abi.ExtMCMCError) as ei:
        GpuSession(em.GsnTargetLaw([0.0]), [em.MALAUpdate(0.1, [1], prior=em.ImproperPosPrior())], x, [0.0, 1.0], 4)
    assert ei.value.code == _abi.EUNSUPPORTED


def test_full_size_cfg4_replay_parity():
    """BASELINE cfg 4 per-GPU size (8 groups x 4096 observations, 8192 chains) for 6 iterations of the
    MALA + RW + RW-pos schedule: the oracle runs chains 0..255, the other chains are replicas."""
    G, ng, Cn, sub, M = 8, 4096, 8192, 256, 6
    y, grp, _ = _hier_data(G, ng, seed=5)
    law = em.HierNormalLaw(G)
    ups = _hier_updates(G, tau=0.03)
    th_sub = _hier_theta0(G, sub)
    steps = list(em.MCMCSchedule(M, 3))
    o = orc.Oracle(law, ups, y, th_sub, sub, seed=3, y=grp)
    ro = o.run(steps, n_threads=8)
    reps = Cn // sub
    g = GpuSession(law, ups, y, np.tile(th_sub, (1, reps)), Cn, seed=3, n_steps_hint=len(steps), y=grp)
    rg = g.run(steps, replay=(np.tile(ro["proposals"], (1, 1, reps)), np.tile(ro["exp_draws"], (1, reps))))
    from tests.parity import compare_histories
    rep = compare_histories(ro, {k: v[..., :sub] for k, v in rg.items() if hasattr(v, "shape")})
    assert rep["accept_mismatch"] == 0 and rep["near_ties"] == 0 and rep["theta_bitexact"], rep
    assert rep["ll_rel_err"] < 1e-10, rep
    full = rg["theta"].reshape(rg["theta"].shape[:-1] + (reps, sub))
    assert np.array_equal(full, np.broadcast_to(full[..., :1, :], full.shape))
    assert np.array_equal(g.eps(1)[:, :sub], o.eps(1))
    g.close()


def _hier_own_stream(cache, use_graphs, block, n_chains=130, n_iters=36):
    """cfg 4 schedule plus a fourth update (a random walk on theta_1, which DOES move the data term)
    with exclusions, on the device's own Philox stream; EXTMCMC_DATA_CACHE is read when the sweep is
    planned, i.e. per handle."""
    import os
    G = 5
    y, grp, _ = _hier_data(G, 300, seed=21, ragged=True)
    ups = _hier_updates(G) + [em.RandomWalkUpdate(em.UniformRandomWalk([0.2]), [1])]
    steps = list(em.MCMCSchedule(n_iters, len(ups), [(1, range(4, 7)), (4, range(9, 30, 2)), (3, range(20, 24))]))
    old = os.environ.get("EXTMCMC_DATA_CACHE")
    os.environ["EXTMCMC_DATA_CACHE"] = "1" if cache else "0"
    try:
        g = GpuSession(em.HierNormalLaw(G), ups, y, _hier_theta0(G, n_chains), n_chains, seed=31,
                       n_steps_hint=len(steps), use_graphs=use_graphs, y=grp)
        l0 = g.lib.extmcmc_launch_count(g.h)
        parts = [g.run(steps[b:b + block]) for b in range(0, len(steps), block)]
        launches = int(g.lib.extmcmc_launch_count(g.h) - l0)
    finally:
        if old is None:
            del os.environ["EXTMCMC_DATA_CACHE"]
        else:
            os.environ["EXTMCMC_DATA_CACHE"] = old
    assert all(q["rc"] == 0 for q in parts)
    out = {k: np.concatenate([q[k] for q in parts]) for k in ("theta", "theta_prop", "ll", "accepted")}
    out["eps"] = [g.eps(u + 1) for u in range(len(ups))]
    out["stats"] = g.stats()
    out["launches"] = launches
    g.close()
    return out


def test_data_sum_cache_reuses_the_sums_and_changes_nothing_else():
    """HIER_NORMAL: elements that move only mu / tau run without a sweep, a MALA element after them
    finishes its current-state gradient from the cached per-group sums.  Same proposals, decisions,
    trajectories, step sizes and moments as with every element sweeping (the log-likelihood differs only
    by the association of the observation sums), whatever the block length, graphs or not."""
    ref = _hier_own_stream(cache=False, use_graphs=0, block=1000)
    for use_graphs, block in ((0, 1000), (1, 7), (1, 4), (0, 1)):
        got = _hier_own_stream(cache=True, use_graphs=use_graphs, block=block)
        assert np.array_equal(got["accepted"], ref["accepted"]), (use_graphs, block)
        assert np.array_equal(got["theta"], ref["theta"]) and np.array_equal(got["theta_prop"], ref["theta_prop"])
        assert np.allclose(got["ll"][1:], ref["ll"][1:], rtol=1e-12, atol=0)
        assert all(np.array_equal(p, q) for p, q in zip(got["eps"], ref["eps"]))
        for k in ("mean", "cov", "n_accept", "n_prop"):
            assert np.array_equal(got["stats"][k], ref["stats"][k]), k
        assert block < 1000 or got["launches"] < ref["launches"]
    # the two cached runs with different block lengths are bit-identical, log-likelihoods included
    a = _hier_own_stream(cache=True, use_graphs=1, block=7)
    b = _hier_own_stream(cache=True, use_graphs=0, block=1000)
    assert np.array_equal(a["ll"], b["ll"])


def test_checkpoint_resume_hier_with_the_data_sum_cache():
    """The data-sum cache is not part of the checkpoint blob: a handle that resumes in the middle of an
    iteration -- right after the MALA element, in front of the mu / tau elements that would have read
    the cache -- refills it with the kernels that produced it and continues bit for bit, log-likelihoods
    included.  The same for a schedule that starts with a data-free element."""
    import ctypes as C
    G, n_chains = 5, 90
    y, grp, _ = _hier_data(G, 200, seed=4)
    ups = _hier_updates(G)
    th0 = _hier_theta0(G, n_chains)
    steps = list(em.MCMCSchedule(20, len(ups), [(1, range(1, 3))]))   # the first two iterations start at mu
    mk = lambda th: GpuSession(em.HierNormalLaw(G), _hier_updates(G), y, th, n_chains, seed=17,
                               n_steps_hint=len(steps), y=grp, roll_window=10)
    full = mk(th0)
    rf = full.run(steps)
    cut = next(i for i, s in enumerate(steps) if s.mcmciter == 9 and s.pidx == 1) + 1   # after the MALA element
    a = mk(th0)
    ra = a.run(steps[:cut])
    n = C.c_int64()
    a.ck(a.lib.extmcmc_checkpoint_size(a.h, C.byref(n)))
    blob = (C.c_uint8 * n.value)()
    a.ck(a.lib.extmcmc_checkpoint_save(a.h, blob, n.value))
    a.close()
    b = mk(th0 * 0 + 1.0)
    b.ck(b.lib.extmcmc_checkpoint_load(b.h, blob, n.value))
    b.seq = cut
    rb = b.run(steps[cut:])
    for k in ("theta", "theta_prop", "ll", "accepted"):
        assert np.array_equal(rf[k], np.concatenate([ra[k], rb[k]])), k
    sf, sb = full.stats(), b.stats()
    for k in ("mean", "cov", "rolling_ar", "n_accept", "n_prop"):
        assert np.array_equal(sf[k], sb[k]), k
    assert all(np.array_equal(full.eps(u + 1), b.eps(u + 1)) for u in range(len(ups)))
    full.close(); b.close()
