"""GaussianRandomWalk / GaussianRandomWalkMix (random_walk.jl:123-232) and HaarioTypeAdaptation
(adaptation.jl:372-426) on the GPU against the oracle."""
import math

import numpy as np
import pytest

import extensiblemcmc_jl_b200 as em
from extensiblemcmc_jl_b200 import _abi
from oracle import oracle as orc
from tests.parity import GpuSession, replay_compare, theta_init_for

pytestmark = pytest.mark.gpu


def _x(n, seed=0):
    return 1.5 + 2.0 * np.random.default_rng(seed).standard_normal(n)


def _clean(rep):
    assert rep["accept_mismatch"] == 0 and rep["near_ties"] == 0, rep
    assert rep["theta_bitexact"] and rep["ll_rel_err"] < 1e-10, rep
    assert rep["mean_bitexact"] and rep["cov_bitexact"] and rep["counts_equal"] and rep["final_state_bitexact"], rep
    assert 0.02 < rep["accept_rate"] < 0.98, rep


def test_gaussian_walk_joint_with_positive_coordinate():
    x = _x(2000, 1)
    S = np.array([[0.004, 0.001], [0.001, 0.002]])
    ups = [em.RandomWalkUpdate(em.GaussianRandomWalk(S, [False, True]), [1, 2], prior=em.ImproperPosPrior())]
    rep = replay_compare(x, 80, 60, seed=3, updates=ups)
    _clean(rep)


def test_mixture_walk_with_haario_and_a_second_update():
    # update 1: Gaussian mixture + Haario on mu; update 2: uniform multiplicative walk on sigma^2.
    # Haario registers on BOTH updates' steps (adaptation.jl:399-404); update 2 is excluded for a
    # while so that the own-turn counter M and the global registration count N drift apart.
    x = _x(1500, 2)
    ups = [em.RandomWalkUpdate(em.GaussianRandomWalkMix([[0.003]], [[0.02]], 0.3), [1],
                               adpt=em.HaarioTypeAdaptation([0.0], adapt_every_k_steps=7)),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.08], [True]), [2], prior=em.ImproperPosPrior(),
                               adpt=em.AdaptationUnifRW([0.0], adapt_every_k_steps=5, scale=0.01, offset=1.0))]
    rep = replay_compare(x, 64, 60, seed=5, updates=ups, exclude=[(2, range(10, 20))], block=23,
                         history_window=32)
    _clean(rep)
    assert rep["eps_bitexact"], rep          # Sigma_B after several readjust! calls, and eps of update 2


def test_mixture_walk_haario_state_matches_oracle():
    x = _x(1200, 3)
    law = em.GsnTargetLaw([0.0])
    S0 = np.array([[0.004, 0.0], [0.0, 0.004]])
    ups = [em.RandomWalkUpdate(em.GaussianRandomWalkMix(S0, 4 * S0, 0.5, [False, True]), [1, 2],
                               prior=em.ImproperPosPrior(),
                               adpt=em.HaarioTypeAdaptation([0.0, 0.0], adapt_every_k_steps=20))]
    Cn, M = 48, 70
    th0 = theta_init_for(x, Cn)
    steps = list(em.MCMCSchedule(M, 1))
    o = orc.Oracle(law, ups, x, th0, Cn, seed=8)
    ro = o.run(steps, n_threads=8)
    g = GpuSession(law, ups, x, th0, Cn, seed=8, n_steps_hint=M)
    rg = g.run(steps, replay=(ro["proposals"], ro["exp_draws"]))
    assert np.array_equal(ro["accepted"], rg["accepted"]) and np.array_equal(ro["theta"], rg["theta"])
    # the positive coordinate goes through log(): device and host libm may differ in the last ulp
    mo, co = o.adapt_state(1)
    mg, cg = g.adapt_state(1)
    assert np.allclose(mo, mg, rtol=1e-13, atol=0) and np.allclose(co, cg, rtol=1e-9, atol=1e-18)
    assert np.allclose(o.eps(1), g.eps(1), rtol=1e-9, atol=1e-18)
    assert not np.allclose(g.eps(1)[:, 0], (4 * S0).T.ravel())      # Sigma_B was readjusted
    g.close()


def test_tutorial_bivariate_mean_posterior_under_own_philox():
    # docs/src/tutorials/mean_of_bivariate_gaussian.md: posterior of the mean with known Sigma,
    # flat prior: N(xbar, Sigma / n); sampled with the adaptive mixture walk
    rng = np.random.default_rng(0)
    mu, Sig = np.array([1.0, 2.0]), np.array([[1.0, 0.5], [0.5, 1.0]])
    n = 200
    X = rng.multivariate_normal(mu, Sig, size=n)
    law = em.GsnTargetLaw(mu, Sig)
    ups = [em.RandomWalkUpdate(em.GaussianRandomWalkMix(0.01 * np.eye(2), 0.01 * np.eye(2), 0.5), [1, 2],
                               adpt=em.HaarioTypeAdaptation(np.zeros(2), adapt_every_k_steps=50))]
    Cn, M = 128, 3000
    th0 = law.theta.copy(); th0[:2] = 0.0
    mcmc = em.MCMC(ups, backend=em.CUDAMCMCBackend(n_chains=Cn, seed=21, block_len=250))
    ws, lws = em.run_(mcmc, M, dict(P=law, obs=X), th0)
    tr = ws.sub_ws.state_history[1200:, 0, :2]                    # [iters, 2, C]
    ess = em.ess_geyer(tr)
    for k in range(2):
        mcse = math.sqrt(tr[:, k].var() / ess[k].sum())
        assert abs(tr[:, k].mean() - X.mean(axis=0)[k]) < 3 * mcse
        assert abs(tr[:, k].var() - Sig[k, k] / n) < 0.1 * Sig[k, k] / n
    cov01 = np.mean((tr[:, 0] - tr[:, 0].mean()) * (tr[:, 1] - tr[:, 1].mean()))
    assert abs(cov01 - Sig[0, 1] / n) < 0.15 * Sig[0, 1] / n
    acc = ws.stats()["n_accept"].sum() / ws.stats()["n_prop"].sum()
    assert 0.1 < acc < 0.6
    sb = ws.eps(1)                                                # adapted Sigma_B ~ 2.38^2/2 * running cov
    assert sb.shape == (4, Cn) and (sb[0] > 0).all() and (sb[3] > 0).all()
    ws.close()


def test_unsupported_pairings_raise():
    x = _x(50)
    bad = [em.RandomWalkUpdate(em.GaussianRandomWalk([[0.1]]), [1], adpt=em.AdaptationUnifRW([0.0])),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.1]), [1], adpt=em.HaarioTypeAdaptation([0.0]))]
    for u in bad:
        with pytest.raises(_abi.ExtMCMCError) as ei:
            GpuSession(em.GsnTargetLaw([0.0]), [u], x, [0.0, 1.0], 4)
        assert ei.value.code == _abi.EUNSUPPORTED
    # a user closure f(lambda, N, iter) is served by a host callback (extmcmc_set_lambda_fn)
    assert em.HaarioTypeAdaptation([0.0], f=lambda lam, n, it: 0.5 * lam).lambda_callback() is not None
    assert em.HaarioTypeAdaptation([0.0]).lambda_callback() is None


def test_singular_adapted_covariance_is_a_domain_error():
    # k = 1: after the first own turn Sigma_B = 2.38^2 * cov of {0, theta_1} for ONE coordinate is
    # fine, but for two perfectly correlated coordinates it is singular -> MvNormal would throw
    x = _x(200, 4)
    S0 = 1e-4 * np.eye(2)
    ups = [em.RandomWalkUpdate(em.GaussianRandomWalkMix(S0, S0, 1.0), [1, 2],
                               adpt=em.HaarioTypeAdaptation([0.0, 0.0], adapt_every_k_steps=1))]
    g = GpuSession(em.GsnTargetLaw([0.0]), ups, x, [1.5, 4.0], 16, seed=2)
    r = g.run(list(em.MCMCSchedule(4, 1)))
    assert r["rc"] == _abi.EDOMAIN
    g.close()
