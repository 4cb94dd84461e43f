"""The persistent block kernels (csrc/block_kernels.cu): a whole block of schedule elements in one
launch.  Same bars as the per-step path: decisions, trajectories, eps, running moments bit-exact
against the CPU oracle under replayed randomness, log-likelihood within 1e-10 relative; and under
the GPU's own Philox stream the block kernels and the per-step kernels take the same decisions."""
import numpy as np
import pytest

import extensiblemcmc_jl_b200 as em
from extensiblemcmc_jl_b200 import _abi
from tests.parity import GpuSession, cfg2_updates, replay_compare, theta_init_for
from tests.test_gpu_mala import _hier_data, _hier_theta0, _hier_updates

pytestmark = pytest.mark.gpu
RESIDENT, OBS_BLOCK, PER_STEP_CHAINS, PER_STEP_OBS = 3, 4, 1, 2


def _data(n, seed=0, mean=1.5, sd=2.0):
    return mean + sd * np.random.default_rng(seed).standard_normal(n)


def _clean(rep, rate=True):
    assert rep["accept_mismatch"] == 0 and rep["near_ties"] == 0, rep
    assert rep["theta_bitexact"] and rep["ll_rel_err"] < 1e-10, rep
    for k in ("eps_bitexact", "mean_bitexact", "cov_bitexact", "rolling_ar_bitexact", "counts_equal",
              "final_state_bitexact"):
        assert rep[k], (k, rep)
    if rate:
        assert 0.02 < rep["accept_rate"] < 0.98, rep


@pytest.mark.parametrize("n_chains,n_obs,n_iters,block,force,variant", [
    (64, 10000, 40, None, RESIDENT, "team_block_R"),    # 8 chains per CTA: one chain group of 4 per thread group
    (37, 3001, 30, 7, RESIDENT, "team_block_R"),         # odd everything: ragged CTAs, odd N, blocks of 7 elements
    (9, 2050, 30, 1, RESIDENT, "team_block_R"),          # a CTA whose second group holds a single chain; 1-element blocks
    (1, 4097, 40, None, RESIDENT, "team_block_R"),      # one chain: the second thread group is empty
    (1500, 6000, 10, None, RESIDENT, "team_block_R"),    # all 296 CTAs, teams of 4
    (4200, 20000, 6, 4, RESIDENT, "team_block_R"),       # 56-57 chains per team (the cfg 2 layout)
])
def test_resident_replay_parity_gsn1d(n_chains, n_obs, n_iters, block, force, variant):
    rep = replay_compare(_data(n_obs, seed=n_chains), n_chains, n_iters, seed=n_chains + 1, block=block,
                         sweep_variant=force, history_window=max(2 * n_iters, 8))
    assert rep["variant"].startswith(variant), rep["variant"]
    _clean(rep)


@pytest.mark.parametrize("n_obs", [1, 2, 3, 17])
def test_resident_tiny_datasets(n_obs):
    th0 = np.repeat(np.array([[1.0], [2.0]]), 40, axis=1)
    ups = cfg2_updates(eps0=0.8, scale=0.05, k=7, offset=1.0)
    rep = replay_compare(_data(n_obs, seed=9), 40, 50, seed=5, updates=ups, theta_init=th0, sweep_variant=RESIDENT)
    assert rep["variant"].startswith("team_block_R")
    _clean(rep)


def test_resident_exclusions_blocks_and_priors():
    x = _data(4000, seed=4)
    excl = [(1, range(3, 9)), (2, range(4, 30, 2))]
    rep = replay_compare(x, 96, 130, seed=21, exclude=excl, block=37, history_window=40, roll_window=20,
                         sweep_variant=RESIDENT)
    _clean(rep)
    ups = [em.RandomWalkUpdate(em.UniformRandomWalk([0.2]), [1], prior=em.StandardPrior(em.Normal(0.0, 5.0))),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.3], [True]), [2], prior=em.StandardPrior(em.Gamma(2.0, 3.0)),
                               adpt=em.AdaptationUnifRW([0.0], adapt_every_k_steps=9, scale=0.02, offset=1.0))]
    rep = replay_compare(x, 50, 60, seed=22, updates=ups, sweep_variant=RESIDENT)
    _clean(rep)


@pytest.mark.parametrize("n_chains,ragged,force", [(200, False, RESIDENT), (7, True, RESIDENT), (1300, False, RESIDENT)])
def test_resident_cfg4_schedule_replay_parity(n_chains, ragged, force):
    # BASELINE cfg 4: MALA on theta_1..8 (gradient sweeps), uniform walk on mu, multiplicative walk on
    # tau; full 10 x 10 covariance (cooperative update inside the thread group)
    G = 8
    y, grp, _ = _hier_data(G, 256, ragged=ragged)
    rep = replay_compare(y, n_chains, 24, seed=12, updates=_hier_updates(G), law=em.HierNormalLaw(G), y=grp,
                         theta_init=_hier_theta0(G, n_chains), exclude=[(2, range(5, 9))], block=31,
                         history_window=80, sweep_variant=force)
    assert rep["variant"].startswith("team_block_R"), rep["variant"]
    _clean(rep)


def test_resident_mala_gsn1d_and_variances_only():
    x = _data(2500, seed=3)
    ups = [em.MALAUpdate(0.05, [1], prior=em.StandardPrior(em.Normal(0.0, 10.0)),
                         adpt=em.AdaptationMALA(adapt_every_k_steps=5, scale=0.003, offset=1.0)),
           em.RandomWalkUpdate(em.UniformRandomWalk([0.05], [True]), [2], prior=em.ImproperPosPrior())]
    rep = replay_compare(x, 90, 50, seed=6, updates=ups, block=17, history_window=40, sweep_variant=RESIDENT)
    _clean(rep)
    G = 8
    y, grp, _ = _hier_data(G, 256)
    rep = replay_compare(y, 150, 20, seed=13, updates=_hier_updates(G), law=em.HierNormalLaw(G), y=grp,
                         theta_init=_hier_theta0(G, 150), history_window=64, stats_mode=1, sweep_variant=RESIDENT)
    _clean(rep)


@pytest.mark.parametrize("n_chains,n_obs,n_iters,block,force,variant", [
    (5, 20001, 60, None, OBS_BLOCK, "obs_block_C8"),   # a handful of chains
    (1, 4097, 120, 9, OBS_BLOCK, "obs_block_C1"),      # the reference's shape: one chain; 9-element blocks
    (20, 9000, 40, 1, OBS_BLOCK, "obs_block_C32"),     # 1-element blocks: every element through its own launch
    (8, 700001, 12, None, OBS_BLOCK, "obs_block_C8"),  # many tiles per segment: the ring wraps across steps
    (3, 5, 30, None, OBS_BLOCK, "obs_block_C4"),       # fewer observations than CTAs
])
def test_obs_block_replay_parity(n_chains, n_obs, n_iters, block, force, variant):
    rep = replay_compare(_data(n_obs, seed=n_chains), n_chains, n_iters, seed=n_chains + 1, block=block,
                         sweep_variant=force, history_window=max(2 * n_iters, 8))
    assert rep["variant"] == variant
    _clean(rep, rate=n_obs > 100)     # (with 5 observations every small move is accepted)


def _own_stream(x, n_chains, n_iters, variant, block, law=None, ups=None, th0=None, y=None):
    law = law or em.GsnTargetLaw([0.0])
    ups = ups or cfg2_updates(eps0=0.05, scale=5e-3, k=10, offset=2.0)
    th0 = th0 if th0 is not None else theta_init_for(x, n_chains)
    steps = list(em.MCMCSchedule(n_iters, len(ups)))
    g = GpuSession(law, ups, x, th0, n_chains, seed=77, n_steps_hint=len(steps), sweep_variant=variant, y=y)
    parts = [g.run(steps[b:b + block]) for b in range(0, len(steps), block)]
    out = {k: np.concatenate([q[k] for q in parts]) for k in ("theta", "theta_prop", "ll", "accepted")}
    out["eps"] = [g.eps(u + 1) for u in range(len(ups))]
    out["name"] = g.variant()
    g.close()
    return out


def test_block_kernels_and_per_step_kernels_take_the_same_decisions():
    """Own Philox stream, no replay: the persistent kernels and the per-step kernels produce the same
    proposals (bit for bit: same stream, same arithmetic) and the same decisions; log-likelihoods
    differ only by the association of the observation sums."""
    x = _data(30000, seed=2)
    a = _own_stream(x, 96, 40, PER_STEP_CHAINS, 80)
    b = _own_stream(x, 96, 40, RESIDENT, 11)
    assert a["name"].startswith("gsn1d_chains") and b["name"].startswith("team_block_R")
    assert np.array_equal(a["accepted"], b["accepted"])
    assert np.array_equal(a["theta"], b["theta"]) and np.array_equal(a["theta_prop"], b["theta_prop"])
    assert np.allclose(a["ll"][1:], b["ll"][1:], rtol=1e-12, atol=0)
    assert all(np.array_equal(p, q) for p, q in zip(a["eps"], b["eps"]))
    c = _own_stream(x, 6, 40, PER_STEP_OBS, 80)
    d = _own_stream(x, 6, 40, OBS_BLOCK, 7)
    assert c["name"] == "gsn1d_obs_C8" and d["name"] == "obs_block_C8"
    assert np.array_equal(c["accepted"], d["accepted"]) and np.array_equal(c["theta"], d["theta"])
    assert np.allclose(c["ll"][1:], d["ll"][1:], rtol=1e-12, atol=0)
    G = 8
    y, grp, _ = _hier_data(G, 300, seed=4)
    kw = dict(law=em.HierNormalLaw(G), ups=_hier_updates(G), th0=_hier_theta0(G, 120), y=grp)
    e = _own_stream(y, 120, 20, PER_STEP_CHAINS, 60, **kw)
    f = _own_stream(y, 120, 20, RESIDENT, 13, **kw)
    assert f["name"].startswith("team_block_R")
    # MALA proposals use the gradient, i.e. the observation sums: same decisions, states equal up to
    # the association of those sums
    assert np.array_equal(e["accepted"], f["accepted"]) and np.allclose(e["theta"], f["theta"], rtol=1e-9, atol=1e-12)


def test_full_size_cfg2_resident_block():
    """BASELINE cfg 2 shape (4096 chains, N = 1e6) for a few iterations through the resident kernel:
    the oracle runs chains 0..63, the other chains are replicas of them."""
    from oracle import oracle as orc
    from tests.parity import compare_histories
    Cn, sub, M = 4096, 64, 3
    x = _data(1_000_000, seed=2)
    ups = cfg2_updates(eps0=5e-3, scale=5e-4, k=2, offset=1.0)
    th_sub = theta_init_for(x, sub)
    steps = list(em.MCMCSchedule(M, 2))
    o = orc.Oracle(em.GsnTargetLaw([0.0]), ups, x, th_sub, sub, seed=3)
    ro = o.run(steps, n_threads=8)
    reps = Cn // sub
    g = GpuSession(em.GsnTargetLaw([0.0]), ups, x, np.tile(th_sub, (1, reps)), Cn, seed=3, n_steps_hint=len(steps),
                   sweep_variant=RESIDENT)
    assert g.variant() == "team_block_R7"
    rg = g.run(steps, replay=(np.tile(ro["proposals"], (1, 1, reps)), np.tile(ro["exp_draws"], (1, reps))))
    rep = compare_histories(ro, {k: v[..., :sub] for k, v in rg.items() if hasattr(v, "shape")})
    assert rep["accept_mismatch"] == 0 and rep["near_ties"] == 0 and rep["theta_bitexact"], rep
    assert rep["ll_rel_err"] < 1e-10, rep
    full = rg["theta"].reshape(rg["theta"].shape[:-1] + (reps, sub))
    assert np.array_equal(full, np.broadcast_to(full[..., :1, :], full.shape))
    assert np.array_equal(g.eps(1)[:, :sub], o.eps(1)) and np.array_equal(g.eps(2)[:, :sub], o.eps(2))
    g.close()
