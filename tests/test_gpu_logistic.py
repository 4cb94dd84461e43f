"""Bayesian logistic regression (BASELINE cfg 3) on the FP64 tensor-core path against the oracle.
No reference law exists; pinned by scipy/numpy closed forms, finite differences and posterior
agreement with an independent oracle run."""
import math

import numpy as np
import pytest

import extensiblemcmc_jl_b200 as em
from extensiblemcmc_jl_b200 import _abi
from oracle import oracle as orc
from tests.parity import GpuSession, replay_compare

pytestmark = pytest.mark.gpu


def _data(n, d, seed=4):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)) / math.sqrt(d)
    beta = rng.standard_normal(d)
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-X @ beta))).astype(np.float64)
    return X, y, beta


def _np_ll_grad(X, y, th):
    z = X @ th                                          # [n, C]
    ll = (y[:, None] * z - np.logaddexp(0.0, z)).sum(axis=0)
    g = X.T @ (y[:, None] - 1.0 / (1.0 + np.exp(-z)))
    return ll, g


# the last three exercise the flattened work split: more chain blocks than SMs (a CTA walks several
# whole blocks), a single tile per block, and one block whose tiles are spread over all CTAs
@pytest.mark.parametrize("n,d,n_chains", [(1000, 8, 64), (777, 20, 100), (2049, 64, 130), (600, 256, 70), (50, 3, 5),
                                          (100, 8, 64 * 150 + 5), (10, 5, 700), (5000, 12, 33)])
def test_loglik_and_gradient(n, d, n_chains):
    X, y, beta = _data(n, d)
    law = em.LogisticLaw(d)
    rng = np.random.default_rng(1)
    th0 = beta[:, None] + 0.3 * rng.standard_normal((d, n_chains))
    ups = [em.MALAUpdate(0.01, list(range(1, d + 1)))]
    g = GpuSession(law, ups, X, th0, n_chains, y=y)
    assert g.variant() == "logistic_dmma"
    ll, gr = g.eval_grad()
    llw, grw = _np_ll_grad(X, y, th0)
    assert np.allclose(ll, llw, rtol=1e-11, atol=0)
    assert np.allclose(gr, grw, rtol=1e-9, atol=1e-10)
    assert np.allclose(g.eval_loglik(), llw, rtol=1e-11, atol=0)
    o = orc.Oracle(law, ups, X, th0[:, :4], 4, y=y)
    llo, gro = o.loglik_grad(th0[:, :4])
    assert np.allclose(ll[:4], llo, rtol=1e-11, atol=0) and np.allclose(gr[:, :4], gro, rtol=1e-9, atol=1e-10)
    g.close()


def test_mala_replay_parity_logistic():
    d = 16
    X, y, beta = _data(1500, d, seed=9)
    law = em.LogisticLaw(d)
    rng = np.random.default_rng(2)
    Cn = 96
    th0 = 0.01 * rng.standard_normal((d, Cn))
    ups = [em.MALAUpdate(0.15, list(range(1, d + 1)), prior=em.StandardPrior(em.Normal(0.0, 10.0)),
                         adpt=em.AdaptationMALA(adapt_every_k_steps=5, scale=0.01, offset=1.0))]
    rep = replay_compare(X, Cn, 40, seed=3, updates=ups, law=law, y=y, theta_init=th0, block=13,
                         history_window=32)
    assert rep["variant"] == "logistic_dmma"
    assert rep["accept_mismatch"] == 0 and rep["near_ties"] == 0, rep
    assert rep["theta_bitexact"] and rep["ll_rel_err"] < 1e-10, rep
    assert rep["eps_bitexact"] and rep["counts_equal"] and rep["final_state_bitexact"], rep
    assert 0.05 < rep["accept_rate"] < 0.98


def test_random_walk_on_the_logistic_law():
    d = 4
    X, y, beta = _data(800, d, seed=7)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.2] * 2), [1, 2]),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.2] * 2), [3, 4], prior=em.StandardPrior(em.Normal(0.0, 5.0)))]
    rep = replay_compare(X, 70, 30, seed=5, updates=ups, law=em.LogisticLaw(d), y=y,
                         theta_init=np.zeros((d, 70)))
    assert rep["accept_mismatch"] == 0 and rep["theta_bitexact"] and rep["ll_rel_err"] < 1e-10, rep
    assert rep["mean_bitexact"] and rep["cov_bitexact"], rep


def test_posterior_matches_oracle_under_own_philox():
    d = 6
    X, y, beta = _data(400, d, seed=12)
    law = em.LogisticLaw(d)
    mk = lambda: [em.MALAUpdate(0.3, list(range(1, d + 1)), prior=em.StandardPrior(em.Normal(0.0, 10.0)),
                                adpt=em.AdaptationMALA(adapt_every_k_steps=25, scale=0.02, offset=2.0))]
    Cn, M = 256, 1200
    mcmc = em.MCMC(mk(), backend=em.CUDAMCMCBackend(n_chains=Cn, seed=5, block_len=100))
    ws, _ = em.run_(mcmc, M, dict(P=law, obs=X, y=y), np.zeros(d))
    tr = ws.sub_ws.state_history[400:, 0]
    o = orc.Oracle(law, mk(), X, np.zeros(d), 32, seed=99, y=y)
    ro = o.run(list(em.MCMCSchedule(M, 1)), n_threads=8, record=False)
    tro = ro["theta"][400:]
    ess, esso = em.ess_geyer(tr), em.ess_geyer(tro)
    for k in range(d):
        se = math.sqrt(tr[:, k].var() / ess[k].sum() + tro[:, k].var() / esso[k].sum())
        assert abs(tr[:, k].mean() - tro[:, k].mean()) < 3 * se, k
    acc = ws.stats()["n_accept"].sum() / ws.stats()["n_prop"].sum()
    assert 0.35 < acc < 0.8
    ws.close()


def test_full_size_cfg3_loglik_and_gradient():
    """BASELINE cfg 3 at full size (d = 256, N = 1e6, 1024 chains): log-likelihood and gradient of
    every chain in one sweep; 8 chains are checked against the numpy closed form (1e-10)."""
    n, d, Cn = 1_000_000, 256, 1024
    rng = np.random.default_rng(4)
    X = rng.standard_normal((n, d)) / math.sqrt(d)
    beta = rng.standard_normal(d)
    y = (rng.random(n) < 1.0 / (1.0 + np.exp(-X @ beta))).astype(np.float64)
    th0 = beta[:, None] + 0.05 * rng.standard_normal((d, Cn))
    g = GpuSession(em.LogisticLaw(d), [em.MALAUpdate(0.02, list(range(1, d + 1)))], X, th0, Cn, y=y)
    ll, gr = g.eval_grad()
    sub = [0, 1, 63, 64, 500, 777, 1022, 1023]
    llw, grw = _np_ll_grad(X, y, th0[:, sub])
    assert np.allclose(ll[sub], llw, rtol=1e-10, atol=0)
    assert np.allclose(gr[:, sub], grw, rtol=1e-9, atol=1e-8)
    assert np.isfinite(ll).all() and np.isfinite(gr).all()
    g.close()
