"""General-d GsnTargetLaw (theta = [mu; vec(Sigma)], src/example/gsn_target.jl:1-29) on the GPU
against the oracle and scipy."""
import numpy as np
import pytest
from scipy import stats

import extensiblemcmc_jl_b200 as em
from oracle import oracle as orc
from tests.parity import GpuSession, replay_compare

pytestmark = pytest.mark.gpu


def _mv_data(d, n, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d, d))
    Sig = A @ A.T / d + np.eye(d)
    mu = rng.standard_normal(d)
    return mu, Sig, rng.multivariate_normal(mu, Sig, size=n)


def _theta_init(law, n_chains, seed=1):
    rng = np.random.default_rng(seed)
    d = law.d
    th = np.repeat(law.theta[:, None], n_chains, axis=1)
    th[:d] += 0.05 * rng.standard_normal((d, n_chains))
    for k in range(d):                                  # jitter the variances, keep SPD
        th[d + k + k * d] *= np.exp(0.05 * rng.standard_normal(n_chains))
    return th


def _updates(d):
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.15]), [k + 1],
                               adpt=em.AdaptationUnifRW([0.0], adapt_every_k_steps=8, scale=0.01, offset=1.0))
           for k in range(d)]
    # variances: multiplicative walk with the 1/x prior; one covariance entry (upper triangle): additive
    ups.append(em.RandomWalkUpdate(em.UniformRandomWalk([0.1], [True]), [d + 1], prior=em.ImproperPosPrior()))
    ups.append(em.RandomWalkUpdate(em.UniformRandomWalk([0.02]), [d + 1 + d]))          # Sigma[1, 2]
    return ups


@pytest.mark.parametrize("d,n_chains,n_obs,variant", [
    (2, 96, 1501, "gsnmv_chains"),
    (3, 200, 800, "gsnmv_chains"),
    (2, 1, 10, "gsnmv_obs"),            # the reference's own test shape (test/runtests.jl:87-114)
    (4, 6, 3000, "gsnmv_obs"),
    (8, 40, 700, "gsnmv_chains"),
])
def test_replay_parity_general_d(d, n_chains, n_obs, variant):
    mu, Sig, X = _mv_data(d, n_obs, seed=d)
    law = em.GsnTargetLaw(mu, Sig)
    rep = replay_compare(X, n_chains, 30, seed=d + 10, updates=_updates(d), law=law,
                         theta_init=_theta_init(law, n_chains))
    assert rep["variant"] == variant
    assert rep["accept_mismatch"] == 0 and rep["near_ties"] == 0, rep
    assert rep["theta_bitexact"] and rep["ll_rel_err"] < 1e-10, rep
    assert rep["eps_bitexact"] and rep["mean_bitexact"] and rep["cov_bitexact"] and rep["counts_equal"], rep
    assert 0.02 < rep["accept_rate"] < 0.98


@pytest.mark.parametrize("d", [2, 5, 8])
def test_loglik_against_scipy(d):
    mu, Sig, X = _mv_data(d, 5000, seed=20 + d)
    law = em.GsnTargetLaw(mu, Sig)
    Cn = 70
    th0 = _theta_init(law, Cn)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.1]), [1])]
    s = GpuSession(law, ups, X, th0, Cn)
    got = s.eval_loglik()
    for c in (0, 17, 69):
        S = th0[d:, c].reshape(d, d).T            # column-major vec
        S = np.triu(S) + np.triu(S, 1).T          # Symmetric(triu(S))
        want = stats.multivariate_normal.logpdf(X, th0[:d, c], S).sum()
        assert abs(got[c] - want) < 1e-10 * abs(want)
    s.close()


def test_non_spd_covariance_is_a_domain_error():
    mu, Sig, X = _mv_data(2, 50, seed=1)
    law = em.GsnTargetLaw(mu, Sig)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([50.0]), [5])]      # wild walk on Sigma[1, 2]
    s = GpuSession(law, ups, X, law.theta, 32, seed=3)
    r = s.run(list(em.MCMCSchedule(6, 1)))
    from extensiblemcmc_jl_b200 import _abi
    assert r["rc"] == _abi.EDOMAIN
    s.close()
