"""What round 2 lifted or added on the device: general-d GsnTargetLaw up to d = 16, random walks up to
32 coordinates (Gaussian mixture + Haario at p_u = 16), StandardPrior(MvNormal) on a joint update,
HaarioTypeAdaptation's weight schedule f(lambda, N, iter) as a host callback, and the device's OWN
draw arithmetic (Philox mode) for positive coordinates and Gaussian walks against the oracle's."""
import numpy as np
import pytest

import extensiblemcmc_jl_b200 as em
from extensiblemcmc_jl_b200 import _abi
from oracle import oracle as orc
from tests.parity import GpuSession, replay_compare, theta_init_for
from tests.test_gpu_gsnmv import _mv_data, _theta_init, _updates

pytestmark = pytest.mark.gpu


def _clean(rep, eps=True):
    assert rep["accept_mismatch"] == 0 and rep["near_ties"] == 0, rep
    assert rep["theta_bitexact"] and rep["ll_rel_err"] < 1e-10, rep
    assert rep["mean_bitexact"] and rep["cov_bitexact"] and rep["counts_equal"] and rep["final_state_bitexact"], rep
    if eps:
        assert rep["eps_bitexact"], rep
    assert 0.02 < rep["accept_rate"] < 0.98, rep


@pytest.mark.parametrize("d,n_chains,n_obs,variant", [
    (12, 70, 400, "gsnmv_chains"),       # mu and W in shared memory (one column per thread)
    (16, 130, 300, "gsnmv_chains"),      # the oracle's own limit; 3 CTAs of 64 chains, the last ragged
    (16, 5, 700, "gsnmv_obs"),           # observation-mapped: one broadcast column
    (9, 33, 250, "gsnmv_chains"),
])
def test_replay_parity_general_d_up_to_16(d, n_chains, n_obs, variant):
    mu, Sig, X = _mv_data(d, n_obs, seed=d)
    law = em.GsnTargetLaw(mu, Sig)
    rep = replay_compare(X, n_chains, 12, seed=d + 10, updates=_updates(d), law=law,
                         theta_init=_theta_init(law, n_chains), stats_mode=1)
    assert rep["variant"] == variant
    _clean(rep)


def test_loglik_d16_against_scipy():
    from scipy import stats
    d = 16
    mu, Sig, X = _mv_data(d, 2000, seed=36)
    law = em.GsnTargetLaw(mu, Sig)
    th0 = _theta_init(law, 70)
    s = GpuSession(law, [em.RandomWalkUpdate(em.UniformRandomWalk([0.1]), [1])], X, th0, 70, stats_mode=1)
    got = s.eval_loglik()
    for c in (0, 33, 69):
        S = th0[d:, c].reshape(d, d).T
        S = np.triu(S) + np.triu(S, 1).T
        want = stats.multivariate_normal.logpdf(X, th0[:d, c], S).sum()
        assert abs(got[c] - want) < 1e-10 * abs(want)
    s.close()


def test_gaussian_mixture_walk_with_haario_on_16_coordinates():
    # a 16-coordinate joint update of the mean of a d = 16 Gaussian law: Cholesky factors cached per
    # chain and refreshed at every readjust!, running covariance 16 x 16 per chain
    d = 16
    mu, Sig, X = _mv_data(d, 120, seed=5)
    law = em.GsnTargetLaw(mu, Sig)
    SA = 0.002 * (np.eye(d) + 0.1 * np.ones((d, d)))
    ups = [em.RandomWalkUpdate(em.GaussianRandomWalkMix(SA, 2.0 * SA, 0.4), list(range(1, d + 1)),
                               adpt=em.HaarioTypeAdaptation(np.zeros(d), adapt_every_k_steps=60)),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.05], [True]), [d + 1], prior=em.ImproperPosPrior())]
    # (k = 60 own turns: by then the chain has visited more than 16 distinct states, so the adapted
    # 16 x 16 Sigma_B is positive definite; with fewer the reference throws PosDefException)
    rep = replay_compare(X, 24, 120, seed=3, updates=ups, law=law, theta_init=_theta_init(law, 24), stats_mode=1,
                         history_window=256)
    _clean(rep, eps=False)
    assert rep["eps_max_rel"] < 1e-9, rep      # Sigma_B after two readjustments (no positive coordinates)


def test_uniform_walk_on_32_coordinates_and_the_limit():
    d = 5                                       # 30 parameters
    mu, Sig, X = _mv_data(d, 200, seed=2)
    law = em.GsnTargetLaw(mu, Sig)
    coords = list(range(1, 6)) + [6]            # the means and Sigma[1,1] jointly ...
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.05] * 5 + [0.02], [False] * 5 + [True]), coords,
                               prior=em.ProductPrior([em.Normal(0.0, 10.0), em.ImproperPosPrior()], [5, 1]),
                               adpt=em.AdaptationUnifRW([0.0], adapt_every_k_steps=5, scale=0.002, offset=1.0))]
    rep = replay_compare(X, 50, 40, seed=4, updates=ups, law=law, theta_init=_theta_init(law, 50))
    _clean(rep)
    big = em.GsnTargetLaw(np.zeros(6))          # 42 parameters
    for n, ok in ((32, True), (33, False)):
        u = em.RandomWalkUpdate(em.UniformRandomWalk([0.1] * n), list(range(1, n + 1)))
        if ok:
            GpuSession(big, [u], np.zeros((5, 6)), big.theta, 4, stats_mode=1).close()
        else:
            with pytest.raises(_abi.ExtMCMCError) as ei:
                GpuSession(big, [u], np.zeros((5, 6)), big.theta, 4, stats_mode=1)
            assert ei.value.code == _abi.EUNSUPPORTED


def test_mvnormal_prior_on_a_joint_update():
    x = 1.5 + 2.0 * np.random.default_rng(1).standard_normal(600)
    Sg = np.array([[4.0, 1.5], [1.5, 9.0]])
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.2, 0.3]), [1, 2],
                               prior=em.StandardPrior(em.MvNormal([1.0, 5.0], Sg)),
                               adpt=em.AdaptationUnifRW([0.0], adapt_every_k_steps=6, scale=0.01, offset=1.0))]
    rep = replay_compare(x, 90, 50, seed=7, updates=ups)
    _clean(rep)      # (the oracle's MvNormal density has its own second lineage in tests/test_pyref.py)


def test_haario_lambda_schedule_is_called_back_on_the_host():
    x = 1.5 + 2.0 * np.random.default_rng(2).standard_normal(800)
    calls = []

    def f(lam, N, it):
        calls.append((lam, N, it))
        return max(0.05, 0.8 * lam)

    ups = [em.RandomWalkUpdate(em.GaussianRandomWalkMix([[0.003]], [[0.02]], 0.6), [1],
                               adpt=em.HaarioTypeAdaptation([0.0], adapt_every_k_steps=5, f=f)),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.08], [True]), [2], prior=em.ImproperPosPrior())]
    rep = replay_compare(x, 32, 40, seed=6, updates=ups, block=13, history_window=32)
    _clean(rep, eps=False)
    # 40 own turns, k = 5 -> 8 readjustments: the oracle calls f once per chain and readjustment, the
    # device once per readjustment (lambda is shared by all chains), with the same arguments:
    # N = registrations so far + 1 (every executed step of any update registers), iter = mcmciter
    assert len(calls) == 8 * 32 + 8
    args = sorted(set(calls))
    assert len(args) == 8                                           # the same 8 (lambda, N, iter) on both sides
    assert [a[2] for a in sorted(args, key=lambda a: a[2])] == [5, 10, 15, 20, 25, 30, 35, 40]
    assert all(a[1] == 2 * a[2] for a in args)                      # 2 updates per iteration, update 1 first: N = 2 iter - 1 + 1
    lams = sorted({a[0] for a in args}, reverse=True)
    assert lams[0] == 0.6 and abs(lams[1] - 0.48) < 1e-15


def test_device_philox_proposals_match_the_oracle_stream():
    """Own Philox stream (no replay): the first proposal of every chain, computed by the DEVICE's draw
    arithmetic -- theta * exp(U) for positive coordinates, Box-Muller (sincospi) + cached Cholesky
    factor for Gaussian walks -- against the oracle's (cos / sin of 2 pi u): within 2 ulp (uniform walk) / 4 ulp (Gaussian walks)."""
    x = 1.5 + 2.0 * np.random.default_rng(3).standard_normal(300)
    Cn = 200
    th0 = theta_init_for(x, Cn)
    S = 0.01 * np.array([[1.0, 0.4], [0.4, 2.0]])
    cases = [
        [em.RandomWalkUpdate(em.UniformRandomWalk([0.2, 0.1], [False, True]), [1, 2], prior=em.ImproperPosPrior())],
        [em.RandomWalkUpdate(em.GaussianRandomWalk(S, [False, True]), [1, 2], prior=em.ImproperPosPrior())],
        [em.RandomWalkUpdate(em.GaussianRandomWalkMix(S, 3 * S, 0.5, [False, True]), [1, 2], prior=em.ImproperPosPrior())],
    ]
    for ups, max_ulp in zip(cases, (2.0, 4.0, 4.0)):     # exp / log / sincospi vs cos, sin of libm, chained
        steps = list(em.MCMCSchedule(1, 1))
        o = orc.Oracle(em.GsnTargetLaw([0.0]), ups, x, th0, Cn, seed=99, chain_offset=11)
        ro = o.run(steps)
        g = GpuSession(em.GsnTargetLaw([0.0]), ups, x, th0, Cn, seed=99, chain_offset=11, n_steps_hint=1)
        rg = g.run(steps)
        a, b = ro["theta_prop"][0], rg["theta_prop"][0]
        ulp = np.abs(a - b) / np.spacing(np.abs(a))
        assert ulp.max() <= max_ulp, ulp.max()
        assert np.array_equal(ro["accepted"], rg["accepted"])
        g.close()
