"""Pins of the CPU oracle (the reference pins none of the step numerics -- see
oracle/extmcmc_oracle.h): Philox known answers, closed forms against scipy, the analytic
Gaussian mean/variance posterior, and an independent pure-Python restatement of the
transition step replayed on the oracle's own recorded randomness."""
import json
import math
import os

import numpy as np
import pytest
from scipy import stats

import extensiblemcmc_jl_b200 as em
from oracle import oracle as orc
from tests.parity import cfg2_updates

HERE = os.path.dirname(os.path.abspath(__file__))


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors for philox4x32 10 rounds
    assert orc.philox((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert orc.philox((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert orc.philox((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_uniform_stream_layout_and_range():
    seed, chain, it, pidx = 0x1234567890ABCDEF, (5 << 32) | 17, 42, 3
    w = orc.philox((chain & 0xFFFFFFFF, chain >> 32, it, (pidx << 16) | 2), (seed & 0xFFFFFFFF, seed >> 32))
    u4 = (((w[1] << 32) | w[0]) >> 12) + 0.5
    u5 = (((w[3] << 32) | w[2]) >> 12) + 0.5
    assert orc.uniform(seed, chain, it, pidx, 4) == u4 * 2.0 ** -52
    assert orc.uniform(seed, chain, it, pidx, 5) == u5 * 2.0 ** -52
    us = np.array([orc.uniform(9, c, 1, 0, j) for c in range(200) for j in range(10)])
    assert us.min() > 0.0 and us.max() < 1.0
    assert abs(us.mean() - 0.5) < 0.03 and abs(us.var() - 1 / 12) < 0.01


def test_loglik_closed_form_against_scipy():
    rng = np.random.default_rng(5)
    x = 0.3 + 1.7 * rng.standard_normal(257)
    law = em.GsnTargetLaw([0.0])
    o = orc.Oracle(law, cfg2_updates(), x, [0.0, 1.0], n_chains=1)
    th = np.array([[0.1, -2.0, 0.31], [0.5, 3.0, 2.89]])
    got = o.loglik(th)
    want = [stats.norm.logpdf(x, m, math.sqrt(v)).sum() for m, v in th.T]
    assert np.allclose(got, want, rtol=1e-13, atol=0)
    assert np.isnan(o.loglik(np.array([[0.0], [-1.0]]))[0])      # reference: PosDefException


def test_general_d_law_against_scipy():
    rng = np.random.default_rng(6)
    d = 3
    A = rng.standard_normal((d, d))
    Sig = A @ A.T + d * np.eye(d)
    mu = rng.standard_normal(d)
    X = rng.multivariate_normal(mu, Sig, size=50)
    law = em.GsnTargetLaw(mu, Sig)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.1] * d), [1, 2, 3])]
    o = orc.Oracle(law, ups, X, law.theta, n_chains=1)
    got = o.loglik(law.theta[:, None])[0]
    want = stats.multivariate_normal.logpdf(X, mu, Sig).sum()
    assert abs(got - want) <= 1e-12 * abs(want)
    # only the upper triangle of Sigma is read (Symmetric(triu(S)), gsn_target.jl:19)
    th2 = law.theta.copy()
    S2 = Sig.copy(); S2[np.tril_indices(d, -1)] = 123.0
    th2[d:] = S2.T.ravel()
    assert o.loglik(th2[:, None])[0] == got


def _py_step_reference(x, theta, eps, pos_flags, coords, prior_pos, ll_cur, prop_loc, E):
    """Independent pure-Python restatement of one accept/reject (run.jl:268-281)."""
    full = list(theta)
    for c, v in zip(coords, prop_loc):
        full[c] = v
    mu, var = full
    s = math.sqrt(var)
    c0 = -(math.log(2 * math.pi) + 2 * math.log(s)) / 2
    llp = 0.0
    for xi in x:
        z = (xi - mu) / s
        llp += c0 - z * z / 2
    th_loc = [theta[c] for c in coords]
    q = lambda to: sum((-math.log(2.0 * e) - math.log(t)) if f else 0.0 for e, t, f in zip(eps, to, pos_flags))
    lp = lambda th: -sum(math.log(t) for t in th) if prior_pos else 0.0
    llr = llp - ll_cur
    llr = llr + q(th_loc)
    llr = llr - q(prop_loc)
    llr = llr + lp(prop_loc)
    llr = llr - lp(th_loc)
    acc = E > -llr
    return llp, llr, acc, full


def test_three_step_trace_first_accept_reject_and_readjust():
    # exercises: first step ll = -Inf => always accepted (workspaces.jl:425, run.jl:109);
    # a rejection; an eps readjust (k = 2) with delta = scale/sqrt(max(1, iter/k - offset))
    rng = np.random.default_rng(11)
    x = 1.0 + 2.0 * rng.standard_normal(40)
    law = em.GsnTargetLaw([0.0])
    mk = lambda: em.AdaptationUnifRW([0.0], adapt_every_k_steps=2, scale=0.1, offset=0.0)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.5]), [1], adpt=mk()),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.5], [True]), [2], prior=em.ImproperPosPrior(), adpt=mk())]
    steps = list(em.MCMCSchedule(3, 2))
    o = orc.Oracle(law, ups, x, [0.0, 1.0], n_chains=1, seed=3)
    props = np.array([[[0.9]], [[3.5]], [[25.0]], [[0.2]], [[1.1]], [[4.2]]])   # [n][1][1]
    exps = np.array([[0.5], [0.01], [0.3], [0.7], [1e-9], [2.0]])
    r = o.run(steps, replay=(props, exps))
    theta, ll_cur = [0.0, 1.0], -math.inf
    eps = [[0.5], [0.5]]
    counters = [[0, 0], [0, 0]]
    for s, st in enumerate(steps):
        u = st.pidx - 1
        llp, llr, acc, full = _py_step_reference(x, theta, eps[u], [u == 1], [u], u == 1, ll_cur,
                                                 [props[s, 0, 0]], exps[s, 0])
        if s == 0:
            assert llr == math.inf and acc
        assert bool(r["accepted"][s, 0]) == acc
        assert r["ll_prop"][s, 0] == pytest.approx(llp, rel=1e-14)
        if math.isfinite(llr):
            assert r["llr"][s, 0] == pytest.approx(llr, rel=1e-12, abs=1e-12)
        if acc:
            theta, ll_cur = full, llp
        assert list(r["theta"][s, :, 0]) == theta
        assert list(r["theta_prop"][s, :, 0]) == full
        counters[u][0] += 1; counters[u][1] += int(acc)
        if counters[u][0] >= 2:
            delta = 0.1 / math.sqrt(max(1.0, st.mcmciter / 2 - 0.0))
            a_r = counters[u][1] / counters[u][0]
            counters[u] = [0, 0]
            eps[u] = [max(min(eps[u][0] + (1.0 if a_r > 0.234 else -1.0) * delta, 1e7), 1e-12)]
    assert not all(r["accepted"][:, 0])          # the trace contains a rejection
    assert o.eps(1)[0, 0] == eps[0][0] and o.eps(2)[0, 0] == eps[1][0]
    assert o.eps(1)[0, 0] != 0.5


def test_chain_stats_phantom_zero_sample():
    # chain_statistics.jl:46-51 with N starting at 1: mean/cov equal the sample mean and the
    # (n-1)-covariance of {0, theta_1, ..., theta_k}
    rng = np.random.default_rng(2)
    x = rng.standard_normal(30)
    law = em.GsnTargetLaw([0.0])
    o = orc.Oracle(law, cfg2_updates(eps0=0.3), x, [0.0, 1.0], n_chains=2, seed=1)
    steps = list(em.MCMCSchedule(25, 2))
    r = o.run(steps)
    st = o.stats()
    for c in range(2):
        traj = np.vstack([np.zeros((1, 2)), r["theta"][:, :, c]])
        assert np.allclose(st["mean"][:, c], traj.mean(axis=0), rtol=1e-12)
        assert np.allclose(st["cov"][:, :, c], np.cov(traj.T), rtol=1e-10, atol=1e-14)
    assert np.array_equal(st["n_prop"], np.full((2, 2), 25))
    assert np.array_equal(st["n_accept"], r["accepted"].reshape(25, 2, 2).sum(axis=0))


def test_analytic_posterior_cfg1():
    # BASELINE cfg 1: theta = [mu, sigma^2], flat prior on mu, 1/sigma^2 on sigma^2:
    # E[mu] = xbar, Var[mu] = S/(n(n-3)), E[sigma^2] = S/(n-3)
    rng = np.random.default_rng(1)
    n = 1000
    x = 1.0 + 2.0 * rng.standard_normal(n)
    law = em.GsnTargetLaw([0.0])
    mk = lambda: em.AdaptationUnifRW([0.0], adapt_every_k_steps=50, scale=0.1)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.5]), [1], adpt=mk()),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.5], [True]), [2], prior=em.ImproperPosPrior(), adpt=mk())]
    Cn, M = 16, 6000
    o = orc.Oracle(law, ups, x, [0.0, 1.0], n_chains=Cn, seed=2024)
    r = o.run(list(em.MCMCSchedule(M, 2)), n_threads=8, record=False)
    tr = r["theta"].reshape(M, 2, 2, Cn)[2000:, 1]              # post warm-up, after the 2nd update
    S = ((x - x.mean()) ** 2).sum()
    ess = em.ess_geyer(tr)                                      # [2, C]
    for k, (want_mean, want_var) in enumerate([(x.mean(), S / (n * (n - 3))),
                                               (S / (n - 3), 2 * S * S / ((n - 3) ** 2 * (n - 5)))]):
        m = tr[:, k].mean()
        mcse = math.sqrt(tr[:, k].var() / ess[k].sum())
        assert abs(m - want_mean) < 4 * mcse, (k, m, want_mean, mcse)
        v = tr[:, k].var()
        assert abs(v - want_var) < 0.15 * want_var, (k, v, want_var)
    acc = r["accepted"].mean()
    assert 0.15 < acc < 0.35          # adaptation steers towards 0.234


def test_golden_fixture_roundtrip():
    """tests/golden/oracle_cfg_small.json: outputs of the oracle on a fixed input, committed
    with the generating script (tests/golden/make_golden.py); guards against silent drift of
    the restatement and is the fixture the GPU parity test replays."""
    g = json.load(open(os.path.join(HERE, "golden", "oracle_cfg_small.json")))
    x = np.array(g["obs"])
    law = em.GsnTargetLaw([0.0])
    ups = cfg2_updates(**g["update_kwargs"])
    o = orc.Oracle(law, ups, x, np.array(g["theta_init"]), n_chains=g["n_chains"], seed=g["seed"])
    r = o.run(list(em.MCMCSchedule(g["n_iters"], 2)))
    assert np.array_equal(r["accepted"], np.array(g["accepted"], dtype=np.uint8))
    assert np.array_equal(r["theta"], np.array(g["theta"]))
    assert np.allclose(r["ll_prop"], np.array(g["ll_prop"]), rtol=1e-13, atol=0)
    assert np.array_equal(o.eps(1), np.array(g["eps1"])) and np.array_equal(o.eps(2), np.array(g["eps2"]))
