"""GPU parity tests proper: libextmcmc_cuda (through the C ABI) against the CPU oracle on
the same seeded inputs.  Bars: decisions + trajectories bit-exact under replayed
randomness, log-likelihood within 1e-10 relative (tolerance from BASELINE.json), posterior
moments within 3 Monte-Carlo standard errors under the GPU's own Philox stream."""
import json
import math
import os

import numpy as np
import pytest

import extensiblemcmc_jl_b200 as em
from extensiblemcmc_jl_b200 import _abi
from oracle import oracle as orc
from tests.parity import (GpuSession, cfg2_updates, compare_histories, replay_compare,
                          theta_init_for)

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
LL_RTOL = 1e-10


def _data(n, seed=0, mean=1.5, sd=2.0):
    return mean + sd * np.random.default_rng(seed).standard_normal(n)


def _assert_clean(rep):
    assert rep["accept_mismatch"] == 0, rep
    assert rep["near_ties"] == 0, rep           # at these sizes a near-tie would be a bug
    assert rep["theta_bitexact"], rep
    assert rep["ll_rel_err"] < LL_RTOL, rep
    for k in ("eps_bitexact", "mean_bitexact", "cov_bitexact", "rolling_ar_bitexact",
              "counts_equal", "final_state_bitexact"):
        assert rep[k], (k, rep)
    assert 0.02 < rep["accept_rate"] < 0.98, rep


@pytest.mark.parametrize("n_chains,n_obs,n_iters,force,variant", [
    (64, 10000, 60, 0, "gsn1d_chains_R1"),     # BASELINE minimum slice (C=64, N=1e4)
    (300, 3001, 40, 12, "gsn1d_chains_R2"),    # odd N, ragged chain group
    (700, 5000, 25, 14, "gsn1d_chains_R4"),
    (1100, 2050, 25, 18, "gsn1d_chains_R8"),   # 2 chain groups, second one ragged
    (2100, 60000, 12, 0, "gsn1d_chains_R8"),   # what the planner itself picks for a mid-size problem
    (5, 20001, 60, 0, "gsn1d_obs_C8"),         # few chains: observation-mapped kernel
    (1, 4097, 120, 0, "gsn1d_obs_C1"),         # the reference's shape: one chain
    (20, 9000, 40, 0, "gsn1d_obs_C32"),
])
def test_replay_parity(n_chains, n_obs, n_iters, force, variant):
    rep = replay_compare(_data(n_obs, seed=n_chains), n_chains, n_iters, seed=n_chains + 1,
                         sweep_variant=force)
    assert rep["variant"] == variant
    _assert_clean(rep)


@pytest.mark.parametrize("n_obs", [1, 2, 3, 17])
def test_replay_parity_tiny_datasets(n_obs):
    x = _data(n_obs, seed=9)
    th0 = np.repeat(np.array([[1.0], [2.0]]), 40, axis=1)
    ups = cfg2_updates(eps0=0.8, scale=0.05, k=7, offset=1.0)
    rep = replay_compare(x, 40, 50, seed=5, updates=ups, theta_init=th0)
    _assert_clean(rep)


def test_replay_parity_forced_variants_agree():
    x = _data(6000, seed=3)
    for variant in (1, 2):        # the same 48 chains through both kernel mappings
        rep = replay_compare(x, 48, 30, seed=8, sweep_variant=variant)
        _assert_clean(rep)


def test_replay_parity_with_exclusions_and_blocks():
    # excluded updates change the ll hand-over and the rolling acceptance-rate bookkeeping
    x = _data(4000, seed=4)
    excl = [(1, range(3, 9)), (2, range(4, 30, 2))]
    rep = replay_compare(x, 96, 130, seed=21, exclude=excl, block=37, history_window=40,
                         roll_window=20)
    _assert_clean(rep)


def test_replay_no_adaptation_and_priors():
    x = _data(3000, seed=6)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.05]), [1], prior=em.StandardPrior(em.Normal(1.0, 3.0))),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.05], [True]), [2], prior=em.StandardPrior(em.Gamma(2.0, 3.0)))]
    rep = replay_compare(x, 64, 40, seed=2, updates=ups)
    _assert_clean(rep)


def test_replay_further_prior_families():
    # Exponential / InverseGamma / Beta / LogNormal / Cauchy, alone and inside a ProductPrior
    x = _data(2000, seed=9, mean=0.5, sd=0.6)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.05]), [1], prior=em.StandardPrior(em.Beta(2.0, 3.0))),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.08], [True]), [2], prior=em.StandardPrior(em.InverseGamma(3.0, 1.0))),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.04, 0.06], [False, True]), [1, 2],
                               prior=em.ProductPrior([em.Cauchy(0.4, 0.5), em.LogNormal(-1.0, 0.7)], [1, 1])),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.08], [True]), [2], prior=em.StandardPrior(em.Exponential(0.5)))]
    th0 = np.repeat(np.array([[0.5], [0.36]]), 80, axis=1)
    rep = replay_compare(x, 80, 30, seed=11, updates=ups, theta_init=th0)
    _assert_clean(rep)


def test_product_prior_and_support_redraw():
    # ProductPrior over a joint update (priors.jl:60-88); the Uniform factor has bounded support, so
    # the oracle's own stream exercises the whole-vector redraw of proposal! (updates.jl:193-195)
    x = _data(2500, seed=8, mean=0.4, sd=1.0)
    pr = em.ProductPrior([em.Uniform(0.2, 0.6), em.Gamma(2.0, 1.5)], [1, 1])
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.3, 0.08], [False, True]), [1, 2], prior=pr)]
    th0 = np.repeat(np.array([[0.4], [1.0]]), 72, axis=1)
    rep = replay_compare(x, 72, 50, seed=6, updates=ups, theta_init=th0)
    _assert_clean(rep)
    # own Philox stream on the GPU: every proposal and every state stays inside the support
    s = GpuSession(em.GsnTargetLaw([0.0]), ups, x, th0, 72, seed=3, n_steps_hint=50)
    r = s.run(list(em.MCMCSchedule(50, 1)))
    assert (r["theta_prop"][:, 0] >= 0.2).all() and (r["theta_prop"][:, 0] <= 0.6).all()
    assert (r["theta"][:, 0] >= 0.2).all() and (r["theta"][:, 0] <= 0.6).all()
    s.close()


def test_joint_update_of_both_coordinates():
    x = _data(3000, seed=7)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.03, 0.04], [False, True]), [1, 2],
                               prior=em.ImproperPrior(),
                               adpt=em.AdaptationUnifRW([0.0, 0.0], adapt_every_k_steps=6, scale=0.004, offset=1.0))]
    rep = replay_compare(x, 80, 60, seed=4, updates=ups)
    _assert_clean(rep)


def test_golden_fixture_replay_on_gpu():
    g = json.load(open(os.path.join(HERE, "golden", "oracle_cfg_small.json")))
    x = np.array(g["obs"])
    ups = cfg2_updates(**g["update_kwargs"])
    steps = list(em.MCMCSchedule(g["n_iters"], 2))
    s = GpuSession(em.GsnTargetLaw([0.0]), ups, x, np.array(g["theta_init"]), g["n_chains"],
                   seed=g["seed"], n_steps_hint=len(steps))
    r = s.run(steps, replay=(np.array(g["proposals"]), np.array(g["exp_draws"])))
    assert np.array_equal(r["accepted"], np.array(g["accepted"], dtype=np.uint8))
    assert np.array_equal(r["theta"], np.array(g["theta"]))
    assert np.array_equal(r["theta_prop"], np.array(g["theta_prop"]))
    assert np.allclose(r["ll_prop"], np.array(g["ll_prop"]), rtol=LL_RTOL, atol=0)
    assert np.array_equal(s.eps(1), np.array(g["eps1"])) and np.array_equal(s.eps(2), np.array(g["eps2"]))
    s.close()


def test_philox_stream_matches_oracle_bit_for_bit():
    # additive proposals theta + U are exact IEEE arithmetic on both sides, so the first
    # proposal of every chain must agree bit-for-bit if (and only if) the streams agree
    x = _data(500, seed=1)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.3]), [1]),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.7]), [2])]
    th0 = np.repeat(np.array([[1.0], [4.0]]), 333, axis=1)
    steps = list(em.MCMCSchedule(1, 2))[:1]
    for off in (0, (1 << 33) + 5):
        o = orc.Oracle(em.GsnTargetLaw([0.0]), ups, x, th0, 333, seed=0xDEADBEEFCAFE, chain_offset=off)
        ro = o.run(steps)
        g = GpuSession(em.GsnTargetLaw([0.0]), ups, x, th0, 333, seed=0xDEADBEEFCAFE, chain_offset=off)
        rg = g.run(steps)
        assert np.array_equal(ro["theta_prop"], rg["theta_prop"])
        assert np.array_equal(ro["accepted"], rg["accepted"])     # first step: always accepted
        assert ro["accepted"].all()
        g.close()


def test_loglik_full_size_against_sufficient_statistics():
    # size-independent property at BASELINE cfg 2's full size (C = 4096, N = 1e6):
    # sum (x - mu)^2 = Sxx - 2 mu Sx + N mu^2, evaluated in extended precision on the host
    n, Cn = 1_000_000, 4096
    x = _data(n, seed=2)
    th0 = theta_init_for(x, Cn)
    s = GpuSession(em.GsnTargetLaw([0.0]), cfg2_updates(), x, th0, Cn, sweep_variant=1)
    got = s.eval_loglik()
    assert s.variant() == "gsn1d_chains_R8"
    xl = x.astype(np.longdouble)
    Sx, Sxx = xl.sum(), (xl * xl).sum()
    mu, var = th0[0].astype(np.longdouble), th0[1].astype(np.longdouble)
    want = -0.5 * n * np.log(2 * np.pi * var) - (Sxx - 2 * mu * Sx + n * mu * mu) / (2 * var)
    rel = np.abs(got - want.astype(np.float64)) / np.abs(want.astype(np.float64))
    assert rel.max() < LL_RTOL, rel.max()
    # and against the oracle's sequential per-observation sum on a few chains
    o = orc.Oracle(em.GsnTargetLaw([0.0]), cfg2_updates(), x, th0[:, :8], 8)
    ref = o.loglik(th0[:, :8], n_threads=8)
    assert (np.abs(got[:8] - ref) / np.abs(ref)).max() < LL_RTOL
    s.close()


def test_graphs_and_chain_sharding_do_not_change_results():
    x = _data(5000, seed=5)
    Cn = 96
    th0 = theta_init_for(x, Cn)
    ups = cfg2_updates(eps0=0.05, scale=5e-3, k=10, offset=2.0)
    steps = list(em.MCMCSchedule(40, 2))
    law = em.GsnTargetLaw([0.0])
    base = GpuSession(law, ups, x, th0, Cn, seed=77, n_steps_hint=80)
    rb = base.run(steps)
    graph = GpuSession(law, ups, x, th0, Cn, seed=77, n_steps_hint=80, use_graphs=1)
    r1 = graph.run(steps[:50]); r2 = graph.run(steps[50:])
    for k in ("theta", "ll", "accepted", "theta_prop"):
        assert np.array_equal(rb[k], np.concatenate([r1[k], r2[k]]))
    # two "ranks" owning chains [0, 40) and [40, 96): same global Philox key space
    a = GpuSession(law, ups, x, th0[:, :40], 40, seed=77, n_steps_hint=80, chain_offset=0)
    b = GpuSession(law, ups, x, th0[:, 40:], 56, seed=77, n_steps_hint=80, chain_offset=40)
    ra, rbb = a.run(steps), b.run(steps)
    assert np.array_equal(rb["theta"], np.concatenate([ra["theta"], rbb["theta"]], axis=2))
    assert np.array_equal(base.eps(1), np.concatenate([a.eps(1), b.eps(1)], axis=1))
    for s in (base, graph, a, b):
        s.close()


def test_posterior_within_3_mcse_under_own_philox():
    rng = np.random.default_rng(1)
    n = 1000
    x = 1.0 + 2.0 * rng.standard_normal(n)
    mk = lambda: em.AdaptationUnifRW([0.0], adapt_every_k_steps=50, scale=0.1)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.5]), [1], adpt=mk()),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.5], [True]), [2], prior=em.ImproperPosPrior(), adpt=mk())]
    Cn, M = 128, 3000
    mcmc = em.MCMC(ups, backend=em.CUDAMCMCBackend(n_chains=Cn, seed=99, block_len=200))
    ws, lws = em.run_(mcmc, M, dict(P=em.GsnTargetLaw([0.0]), obs=x), [0.0, 1.0])
    tr = ws.sub_ws.state_history[1000:, 1]                    # [iters, p, C] after the 2nd update
    S = ((x - x.mean()) ** 2).sum()
    ess = em.ess_geyer(tr)
    targets = [(x.mean(), S / (n * (n - 3))), (S / (n - 3), 2 * S * S / ((n - 3) ** 2 * (n - 5)))]
    for k, (want_mean, want_var) in enumerate(targets):
        m = tr[:, k].mean()
        mcse = math.sqrt(tr[:, k].var() / ess[k].sum())
        assert abs(m - want_mean) < 3 * mcse, (k, m, want_mean, mcse)
        # variance: MCSE of a variance estimate ~ var * sqrt(2 / ESS)
        v = tr[:, k].var()
        assert abs(v - want_var) < 3 * want_var * math.sqrt(2.0 / ess[k].sum()) + 0.02 * want_var, (k, v, want_var)
    # same model through the oracle's own stream: means agree within 3 combined MCSE
    o = orc.Oracle(em.GsnTargetLaw([0.0]), ups, x, [0.0, 1.0], 32, seed=5)
    ro = o.run(list(em.MCMCSchedule(M, 2)), n_threads=8, record=False)
    tro = ro["theta"].reshape(M, 2, 2, 32)[1000:, 1]
    esso = em.ess_geyer(tro)
    for k in range(2):
        se = math.sqrt(tr[:, k].var() / ess[k].sum() + tro[:, k].var() / esso[k].sum())
        assert abs(tr[:, k].mean() - tro[:, k].mean()) < 3 * se
    st = ws.stats()
    acc = st["n_accept"].sum() / st["n_prop"].sum()
    assert 0.18 < acc < 0.30
    ws.close()


def test_domain_error_is_reported():
    # a free (additive) walk on the variance reaches sigma^2 <= 0: the reference throws
    # (PosDefException); the GPU rejects the proposal and extmcmc_sync returns EDOMAIN
    x = _data(100, seed=1)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([50.0]), [2])]
    s = GpuSession(em.GsnTargetLaw([0.0]), ups, x, [1.5, 4.0], 64, seed=1)
    r = s.run(list(em.MCMCSchedule(5, 1)))
    assert r["rc"] == _abi.EDOMAIN
    assert (r["theta"][:, 1] > 0).all()
    s.close()


def test_unsupported_requests_raise():
    lib = _abi.load()
    cfg = _abi.Config()
    cfg.abi_version, cfg.n_chains, cfg.n_params, cfg.n_updates = _abi.ABI_VERSION, 4, 6, 1
    cfg.law, cfg.obs_dim, cfg.history_window = 99, 2, 4
    import ctypes as C
    h = _abi.Handle()
    assert lib.extmcmc_create(C.byref(cfg), C.byref(h)) == _abi.EUNSUPPORTED
    x = _data(50)
    with pytest.raises(_abi.ExtMCMCError) as ei:
        big = em.GsnTargetLaw(np.zeros(6))      # 42 parameters: a 33-coordinate walk exceeds the device limit (32)
        GpuSession(big, [em.RandomWalkUpdate(em.GaussianRandomWalk(np.eye(33)), list(range(1, 34)))],
                   np.zeros((5, 6)), big.theta, 4)
    assert ei.value.code == _abi.EUNSUPPORTED
    # an empty observation set is refused loudly (EINVAL), as is running before any upload
    cfg.law, cfg.obs_dim, cfg.n_params = _abi.LAW_GSN_IID_1D, 1, 2
    assert lib.extmcmc_create(C.byref(cfg), C.byref(h)) == 0
    empty = np.zeros(1)
    assert lib.extmcmc_upload_obs(h, _abi.dptr(empty), 0, 1, None) == _abi.EINVAL
    step = orc.steps_array(list(em.MCMCSchedule(1, 1)))
    assert lib.extmcmc_run_block(h, step, 1) == _abi.EINVAL
    assert b"observations" in lib.extmcmc_last_error(h)
    lib.extmcmc_destroy(h)


def test_history_ring_staleness():
    x = _data(200, seed=1)
    s = GpuSession(em.GsnTargetLaw([0.0]), cfg2_updates(eps0=0.1), x, [1.5, 4.0], 8, history_window=6)
    steps = list(em.MCMCSchedule(6, 2))
    s.run(steps[:6]); s.run(steps[6:])
    with pytest.raises(_abi.ExtMCMCError) as ei:
        s.history(0, 3)
    assert ei.value.code == _abi.ESTALE
    assert s.history(6, 12)["theta"].shape == (6, 2, 8)
    with pytest.raises(_abi.ExtMCMCError):
        s.run(list(em.MCMCSchedule(7, 1)))          # block longer than the ring
    s.close()


def test_full_size_cfg2_replay_parity():
    """BASELINE cfg 2 at full size (C = 4096, N = 1e6) for 3 iterations.  The oracle runs chains
    0..511 (one chain costs 6e6 logpdf evaluations per iteration); the GPU runs all 4096, chains
    c and c mod 512 being replicas (same start, same replayed randomness), so every chain is
    checked: the first 512 against the oracle, the rest against their replica, bit for bit."""
    n, Cn, sub, M = 1_000_000, 4096, 512, 3
    x = _data(n, seed=2)
    ups = cfg2_updates()                                       # eps0 = 5e-3, k = 50, scale = 5e-4, ...
    th_sub = theta_init_for(x, sub)
    steps = list(em.MCMCSchedule(M, 2))
    o = orc.Oracle(em.GsnTargetLaw([0.0]), ups, x, th_sub, sub, seed=3)
    ro = o.run(steps, n_threads=8)
    reps = Cn // sub
    g = GpuSession(em.GsnTargetLaw([0.0]), ups, x, np.tile(th_sub, (1, reps)), Cn, seed=3, n_steps_hint=len(steps),
                   sweep_variant=1)     # the per-step kernels (the resident block kernel: test_gpu_block_kernels.py)
    assert g.variant() == "gsn1d_chains_R8"
    rg = g.run(steps, replay=(np.tile(ro["proposals"], (1, 1, reps)), np.tile(ro["exp_draws"], (1, reps))))
    first = {k: v[..., :sub] for k, v in rg.items() if hasattr(v, "shape")}
    rep = compare_histories(ro, first)
    assert rep["accept_mismatch"] == 0 and rep["theta_bitexact"], rep
    assert rep["ll_rel_err"] < LL_RTOL, rep                     # observed ~1e-15
    if rep["near_ties"] == 0:
        for k in ("theta", "theta_prop", "ll", "ll_prop", "accepted"):
            full = rg[k].reshape(rg[k].shape[:-1] + (reps, sub))
            assert np.array_equal(full, np.broadcast_to(full[..., :1, :], full.shape)), k
        assert np.array_equal(g.eps(1)[:, :sub], o.eps(1)) and np.array_equal(g.eps(2)[:, :sub], o.eps(2))
    g.close()


def test_device_generated_observations():
    # extmcmc_generate_obs_normal (BASELINE cfg 5 generates its 1e9 observations in place): recover
    # sum x and sum x^2 from log-likelihood evaluations and check them against N(1.5, 2^2); a split
    # into two shards with global offsets must describe the very same data
    import ctypes as C
    lib = _abi.load()
    n = 1_000_001

    def sums(first, count):
        mcmc = em.MCMC(cfg2_updates(), backend=em.CUDAMCMCBackend(n_chains=2, seed=1, history="none"))
        from extensiblemcmc_jl_b200.mcmc import init_
        init_(mcmc, 1, dict(P=em.GsnTargetLaw([0.0]), obs=em.DeviceGeneratedObs(count, 1.5, 2.0, 6, first)),
              np.array([[0.0, 1.0], [1.0, 1.0]]))
        ll = mcmc.workspace.eval_loglik()
        mcmc.workspace.close()
        S = -2.0 * (ll + 0.5 * count * math.log(2 * math.pi))       # S(mu) = sum (x - mu)^2 at var = 1
        sx = (S[0] - S[1] + count) / 2.0
        return sx, S[0]

    sx, sxx = sums(0, n)
    mean, var = sx / n, sxx / n - (sx / n) ** 2
    assert abs(mean - 1.5) < 5 * 2.0 / math.sqrt(n) and abs(var - 4.0) < 5 * 4.0 * math.sqrt(2.0 / n)
    a, b = sums(0, 400_001), sums(400_001, n - 400_001)              # odd split point on purpose
    assert abs((a[0] + b[0]) - sx) < 1e-9 * abs(sx) and abs((a[1] + b[1]) - sxx) < 1e-9 * sxx


def test_many_chains_indexing():
    # 70 001 chains (not a multiple of anything): 64-bit indexing of the SoA arrays, ragged last CTA
    x = _data(600, seed=11)
    Cn = 70_001
    th0 = theta_init_for(x, Cn)
    ups = cfg2_updates(eps0=0.05, scale=5e-3, k=10, offset=2.0)
    steps = list(em.MCMCSchedule(4, 2))
    sub = np.r_[0:40, Cn - 40:Cn]                                   # first and last chains vs the oracle
    o = orc.Oracle(em.GsnTargetLaw([0.0]), ups, x, th0[:, :40], 40, seed=5, chain_offset=0)
    o2 = orc.Oracle(em.GsnTargetLaw([0.0]), ups, x, th0[:, Cn - 40:], 40, seed=5, chain_offset=Cn - 40)
    r1, r2 = o.run(steps), o2.run(steps)
    g = GpuSession(em.GsnTargetLaw([0.0]), ups, x, th0, Cn, seed=5, n_steps_hint=8)
    props = np.zeros((8, 1, Cn)); exps = np.ones((8, Cn))
    props[:, :, sub] = np.concatenate([r1["proposals"], r2["proposals"]], axis=2)
    exps[:, sub] = np.concatenate([r1["exp_draws"], r2["exp_draws"]], axis=1)
    props[:, 0, 40:Cn - 40] = np.where(np.arange(8)[:, None] % 2 == 0, th0[0, 40:Cn - 40], th0[1, 40:Cn - 40])
    rg = g.run(steps, replay=(props, exps))
    ref_theta = np.concatenate([r1["theta"], r2["theta"]], axis=2)
    ref_acc = np.concatenate([r1["accepted"], r2["accepted"]], axis=1)
    assert np.array_equal(rg["theta"][:, :, sub], ref_theta) and np.array_equal(rg["accepted"][:, sub], ref_acc)
    assert np.isfinite(rg["ll"]).all()
    g.close()
