"""The C oracle against a SECOND, independent restatement of the reference (tests/pyref.py: pure
Python, written from the Julia sources with the reference's own object structure).  The oracle runs
under its Philox stream and records proposals and Exp(1) draws; pyref replays chain by chain.
Bars: decisions and trajectories identical, log-likelihoods within 1e-13 relative, step sizes,
running moments and rolling acceptance rates identical (same IEEE operations in the same order).
pyref's own Philox / Box-Muller / proposal arithmetic is checked against the oracle's recorded
proposals as well, so the draw arithmetic has two lineages too."""
import math

import numpy as np
import pytest

import extensiblemcmc_jl_b200 as em
from oracle import oracle as orc
from tests import pyref


def _data(n, seed=0, mean=1.5, sd=2.0):
    return mean + sd * np.random.default_rng(seed).standard_normal(n)


def _replay(law_py, obs, mk_py_updates, res, th0, M, c, exclude=(), roll_window=100, grp=None):
    n_steps = res["accepted"].shape[0]
    draws = pyref.Replay([res["proposals"][k, :, c] for k in range(n_steps)], [res["exp_draws"][k, c] for k in range(n_steps)])
    return pyref.run_chain(law_py, obs, mk_py_updates(), th0[:, c], M, draws, exclude=exclude,
                           roll_window=roll_window, grp=grp)


def _compare(o, res, steps, out, c, NU, ll_rtol=1e-13):
    """res: oracle histories [n_steps, ...]; out: pyref per-(iteration, update) histories."""
    for k, s in enumerate(steps):
        it, pj = s.mcmciter - 1, s.pidx - 1
        assert bool(res["accepted"][k, c]) == bool(out["accepted"][pj][it]), (k, s)
        assert np.array_equal(res["theta"][k, :, c], out["theta"][it, pj]), (k, s)
        assert np.array_equal(res["theta_prop"][k, :, c], out["theta_prop"][it, pj]), (k, s)
        a, b = res["ll_prop"][k, c], out["ll_prop"][pj][it]
        if np.isfinite(a):
            assert abs(a - b) <= ll_rtol * abs(a), (k, a, b)
    st = o.stats()
    assert np.array_equal(st["mean"][:, c], out["mean"])
    assert np.array_equal(st["cov"][:, :, c], out["cov"])
    last = {}
    for s in steps:
        last[s.pidx] = s.mcmciter
    for pj, it in last.items():
        assert st["rolling_ar"][pj - 1, c] == out["rolling_ar"][it - 1][pj - 1]


def test_uniform_walks_adaptation_exclusions_200_iterations():
    x = _data(400, seed=1)
    C, M = 3, 200
    mk = lambda: em.AdaptationUnifRW([0.0], adapt_every_k_steps=7, target_accpt_rate=0.234, scale=0.05, min=1e-6,
                                     max=10.0, offset=2.0)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.3]), [1], adpt=mk(), prior=em.StandardPrior(em.Normal(0.0, 10.0))),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.4], [True]), [2], prior=em.ImproperPosPrior(), adpt=mk())]
    excl = [(1, range(5, 20)), (2, range(30, 90, 3))]
    th0 = np.array([[1.0, 1.4, 2.0], [3.0, 4.0, 5.0]])
    o = orc.Oracle(em.GsnTargetLaw([0.0]), ups, x, th0, C, seed=5, roll_window=20)
    steps = list(em.MCMCSchedule(M, 2, excl))
    res = o.run(steps)
    py_ups = lambda: [pyref.Update("rw", pyref.UniformRW([0.3]), [1], ("std", ("Normal", 0.0, 10.0)),
                                   pyref.AdaptUnifRW(7, 0.234, 0.05, 1e-6, 10.0, 2.0)),
                      pyref.Update("rw", pyref.UniformRW([0.4], [True]), [2], ("improper_pos",),
                                   pyref.AdaptUnifRW(7, 0.234, 0.05, 1e-6, 10.0, 2.0))]
    for c in range(C):
        ups_c = py_ups()
        out = _replay(("gsn", 1), x, lambda: ups_c, res, th0, M, c, exclude=excl, roll_window=20)
        _compare(o, res, steps, out, c, 2)
        for u in (1, 2):
            assert np.array_equal(o.eps(u)[:, c], ups_c[u - 1].rw.eps)
    assert 0.05 < res["accepted"].mean() < 0.95


def test_gaussian_mix_walk_haario_with_lambda_schedule_and_priors():
    d = 2
    rng = np.random.default_rng(3)
    obs = rng.standard_normal((150, d)) @ np.array([[1.0, 0.3], [0.0, 0.8]]) + np.array([0.5, -1.0])
    law = em.GsnTargetLaw(np.zeros(d))
    C, M = 2, 120
    SA, SB = 0.02 * np.array([[1.0, 0.2], [0.2, 1.5]]), 0.05 * np.eye(2)
    f = lambda lam, N, it: 0.5 * lam + 0.25 * (1.0 - 1.0 / (1.0 + N / 50.0))
    ups = [em.RandomWalkUpdate(em.GaussianRandomWalkMix(SA, SB, 0.3), [1, 2],
                               prior=em.StandardPrior(em.MvNormal([0.0, 0.0], [[4.0, 1.0], [1.0, 9.0]])),
                               adpt=em.HaarioTypeAdaptation([0.0, 0.0], adapt_every_k_steps=9, f=f)),
           em.RandomWalkUpdate(em.GaussianRandomWalk(0.01 * np.eye(2), [True, True]), [3, 6],
                               prior=em.ProductPrior([em.Gamma(2.0, 1.5), em.ImproperPosPrior()], [1, 1])),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.05]), [4])]
    th0 = np.repeat(np.array([0.0, 0.0, 1.0, 0.1, 0.1, 1.0])[:, None], C, axis=1) * np.array([1.0, 1.1])[None, :]
    o = orc.Oracle(law, ups, obs, th0, C, seed=9)
    steps = list(em.MCMCSchedule(M, 3, [(3, range(10, 40))]))
    res = o.run(steps)
    L = np.linalg.cholesky(np.array([[4.0, 1.0], [1.0, 9.0]]))
    py_ups = lambda: [pyref.Update("rw", pyref.GaussianRWMix(SA, SB, 0.3), [1, 2], ("mvn", np.zeros(2), L), pyref.Haario(2, 9, f)),
                      pyref.Update("rw", pyref.GaussianRW(0.01 * np.eye(2), [True, True]), [3, 6],
                                   ("product", [(("std", ("Gamma", 2.0, 1.5)), 1), (("improper_pos",), 1)])),
                      pyref.Update("rw", pyref.UniformRW([0.05]), [4])]
    for c in range(C):
        ups_c = py_ups()
        out = _replay(("gsn", d), obs, lambda: ups_c, res, th0, M, c, exclude=[(3, range(10, 40))])
        _compare(o, res, steps, out, c, 3, ll_rtol=1e-12)
        sig_b = o.eps(1)[:, c].reshape(2, 2).T              # column-major Sigma_B after the readjustments
        assert np.array_equal(sig_b, ups_c[0].rw.B.Sigma)
        hm, hc = o.adapt_state(1)
        assert np.array_equal(hm[:, c], ups_c[0].adpt.mean) and np.array_equal(hc[:, c].reshape(2, 2).T, ups_c[0].adpt.cov)
        assert ups_c[0].rw.lam != 0.3                       # the schedule f moved lambda


def test_mala_and_mixed_schedule_on_the_hierarchical_law():
    G = 4
    rng = np.random.default_rng(8)
    tg = rng.standard_normal(G)
    y = np.concatenate([tg[g] + rng.standard_normal(30 + g) for g in range(G)])
    grp = np.concatenate([np.full(30 + g, g) for g in range(G)])
    C, M = 2, 80
    ups = [em.MALAUpdate(0.15, list(range(1, G + 1)), prior=em.StandardPrior(em.Normal(0.0, 5.0)),
                         adpt=em.AdaptationMALA(adapt_every_k_steps=6, scale=0.01, offset=1.0)),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.4]), [G + 1]),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.4], [True]), [G + 2], prior=em.ImproperPosPrior())]
    th0 = np.repeat(np.concatenate([0.1 * np.arange(G), [0.0, 1.0]])[:, None], C, axis=1)
    th0[:, 1] += 0.05
    o = orc.Oracle(em.HierNormalLaw(G), ups, y, th0, C, seed=4, y=grp)
    steps = list(em.MCMCSchedule(M, 3))
    res = o.run(steps)
    py_ups = lambda: [pyref.Update("mala", 0.15, list(range(1, G + 1)), ("std", ("Normal", 0.0, 5.0)),
                                   pyref.AdaptUnifRW(6, 0.574, 0.01, 1e-12, 1e7, 1.0)),
                      pyref.Update("rw", pyref.UniformRW([0.4]), [G + 1]),
                      pyref.Update("rw", pyref.UniformRW([0.4], [True]), [G + 2], ("improper_pos",))]
    for c in range(C):
        ups_c = py_ups()
        out = _replay(("hier", G), y, lambda: ups_c, res, th0, M, c, grp=grp)
        # the hierarchical log-likelihood is summed observation by observation here and group by group
        # in the oracle: values agree to rounding, decisions must still be identical
        _compare(o, res, steps, out, c, 3, ll_rtol=1e-12)
        assert o.eps(1)[0, c] == ups_c[0].tau


def test_own_philox_draws_reproduce_the_recorded_proposals():
    """pyref's Philox4x32-10, uniform mapping, Uniform(-eps, eps) / exp(U) arithmetic and Box-Muller +
    Cholesky proposals against what the oracle recorded (first steps of every chain)."""
    x = _data(50, seed=2)
    C = 5
    S = 0.04 * np.array([[1.0, 0.5], [0.5, 2.0]])
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.3, 0.2], [False, True]), [1, 2]),
           em.RandomWalkUpdate(em.GaussianRandomWalk(S, [False, True]), [1, 2])]
    th0 = np.array([[1.0, 1.2, 1.4, 1.6, 1.8], [3.0, 3.5, 4.0, 4.5, 5.0]])
    seed, off = 0x1234567890ABCDEF, 7
    o = orc.Oracle(em.GsnTargetLaw([0.0]), ups, x, th0, C, seed=seed, chain_offset=off)
    steps = list(em.MCMCSchedule(1, 2))
    res = o.run(steps)
    for c in range(C):
        dr = pyref.PhiloxDraws(seed, off + c)
        dr.start_step(1, 0)
        p0 = pyref.UniformRW([0.3, 0.2], [False, True]).rand(th0[:, c], dr)
        assert np.array_equal(p0, res["proposals"][0, :, c])
        assert dr.exponential() == res["exp_draws"][0, c]
        cur = res["theta"][0, :, c]
        dr.start_step(1, 1)
        p1 = pyref.GaussianRW(S, [False, True]).rand(cur, dr)
        assert np.allclose(p1, res["proposals"][1, :, c], rtol=4e-16, atol=0)     # cos/sin of libm vs Python: <= 2 ulp
        assert dr.exponential() == res["exp_draws"][1, c]


def test_philox_known_answers():
    # Random123 known-answer vectors (also checked against the C oracle in test_oracle.py)
    assert pyref.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert pyref.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert pyref.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
