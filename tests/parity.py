"""Shared parity harness: run the CPU oracle under its own Philox stream, record the
proposals and Exp(1) draws, replay them through libextmcmc_cuda via the C ABI and compare.

Bars (BASELINE.json): accept/reject decisions and trajectories bit-exact under replay
(a decision may differ only at a near-tie |E + llr| <= tie_tol * |ll|, which is reported,
never hidden); log-likelihoods within 1e-10 relative; eps and running moments bit-exact
while the decisions agree.
"""
import ctypes as C

import numpy as np

import extensiblemcmc_jl_b200 as em
from extensiblemcmc_jl_b200 import _abi
from oracle import oracle as orc


def cfg2_updates(eps0=5e-3, scale=5e-4, k=50, vmin=1e-7, offset=100.0):
    mk = lambda: em.AdaptationUnifRW([0.0], adapt_every_k_steps=k, target_accpt_rate=0.234,
                                     scale=scale, min=vmin, max=1e7, offset=offset)
    return [
        em.RandomWalkUpdate(em.UniformRandomWalk([eps0]), [1], adpt=mk()),
        em.RandomWalkUpdate(em.UniformRandomWalk([eps0], [True]), [2],
                            prior=em.ImproperPosPrior(), adpt=mk()),
    ]


def theta_init_for(x, n_chains, seed=3):
    rng = np.random.default_rng(seed)
    th = np.empty((2, n_chains))
    th[0] = x.mean() + 0.01 * rng.standard_normal(n_chains)
    th[1] = x.var(ddof=1) * np.exp(0.01 * rng.standard_normal(n_chains))
    return th


class GpuSession:
    """Thin driver of the C ABI for tests (replay mode lives below the Python mirror)."""

    def __init__(self, law, updates, obs, theta_init, n_chains, seed=0, history_window=None,
                 n_steps_hint=64, y=None, **cfg_kw):
        self.lib = _abi.load()
        self.C, self.p, self.NU = n_chains, law.n_params, len(updates)
        cfg = _abi.Config()
        cfg.abi_version = _abi.ABI_VERSION
        cfg.n_chains, cfg.n_params, cfg.n_updates = n_chains, self.p, self.NU
        cfg.law, cfg.obs_dim, cfg.seed = law.abi_law(), law.obs_dim, seed
        cfg.history_window = history_window or n_steps_hint
        cfg.roll_window = cfg_kw.pop("roll_window", 100)
        cfg.use_graphs = cfg_kw.pop("use_graphs", 0)
        for k, v in cfg_kw.items():
            setattr(cfg, k, v)
        self.stats_mode = int(cfg.stats_mode)
        self.h = _abi.Handle()
        rc = self.lib.extmcmc_create(C.byref(cfg), C.byref(self.h))
        if rc:
            raise _abi.ExtMCMCError(rc, self.lib.extmcmc_last_error(None).decode())
        self._keep = []
        self.p_u = []
        self.kernels = []
        for i, u in enumerate(updates):
            a, keep = u.to_abi(self.p)
            self._keep.append(keep)
            self.p_u.append(len(u.coords))
            self.kernels.append(a.kernel)
            self.ck(self.lib.extmcmc_set_update(self.h, i, C.byref(a)))
            cb = getattr(getattr(u, "adpt", None), "lambda_callback", lambda: None)()
            if cb is not None:                                  # HaarioTypeAdaptation(...; f = ...)
                self._keep.append(cb)
                self.ck(self.lib.extmcmc_set_lambda_fn(self.h, i, cb, None))
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        self.ck(self.lib.extmcmc_upload_obs(self.h, _abi.dptr(obs), obs.shape[0], law.obs_dim, _abi.dptr(y)))
        th = np.asarray(theta_init, dtype=np.float64)
        if th.ndim == 1:
            th = np.repeat(th[:, None], n_chains, axis=1)
        self.ck(self.lib.extmcmc_set_state(self.h, _abi.dptr(np.ascontiguousarray(th))))
        self.seq = 0

    def ck(self, rc):
        return _abi.check(self.h, rc)

    def close(self):
        if self.h:
            self.lib.extmcmc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, steps, replay=None):
        steps = list(steps)
        arr = orc.steps_array(steps)
        n = len(steps)
        if replay is None:
            self.ck(self.lib.extmcmc_run_block(self.h, arr, n))
        else:
            props = np.ascontiguousarray(replay[0], dtype=np.float64)
            exps = np.ascontiguousarray(replay[1], dtype=np.float64)
            self.ck(self.lib.extmcmc_run_block_replay(self.h, arr, n, props.shape[1],
                                                      _abi.dptr(props), _abi.dptr(exps)))
        rc = self.lib.extmcmc_sync(self.h)
        out = self.history(self.seq, self.seq + n)
        out["rc"] = rc
        self.seq += n
        return out

    def history(self, lo, hi):
        n, p, Cn = hi - lo, self.p, self.C
        out = dict(theta=np.empty((n, p, Cn)), theta_prop=np.empty((n, p, Cn)), ll=np.empty((n, Cn)),
                   ll_prop=np.empty((n, Cn)), accepted=np.empty((n, Cn), dtype=np.uint8))
        self.ck(self.lib.extmcmc_get_history(self.h, lo, hi, _abi.dptr(out["theta"]),
                                             _abi.dptr(out["theta_prop"]), _abi.dptr(out["ll"]),
                                             _abi.dptr(out["ll_prop"]),
                                             out["accepted"].ctypes.data_as(_abi.c_uint8_p)))
        return out

    def state(self):
        th, ll = np.empty((self.p, self.C)), np.empty(self.C)
        self.ck(self.lib.extmcmc_get_state(self.h, _abi.dptr(th), _abi.dptr(ll)))
        return th, ll

    def stats(self):
        p, Cn, NU = self.p, self.C, self.NU
        mean, cov, ra = np.empty((p, Cn)), np.empty((p * p, Cn)), np.empty((NU, Cn))
        na, npr = np.empty((NU, Cn), dtype=np.int64), np.empty((NU, Cn), dtype=np.int64)
        self.ck(self.lib.extmcmc_get_stats(self.h, _abi.dptr(mean), _abi.dptr(cov), _abi.dptr(ra),
                                           na.ctypes.data_as(_abi.c_int64_p),
                                           npr.ctypes.data_as(_abi.c_int64_p)))
        # stats_mode 1 keeps the variances only: the library fills the first p rows
        cov = cov[:p] if self.stats_mode == 1 else cov.reshape(p, p, Cn).transpose(1, 0, 2)
        return dict(mean=mean, cov=cov, rolling_ar=ra, n_accept=na, n_prop=npr)

    def eps(self, u):
        n = self.p_u[u - 1]
        k = self.kernels[u - 1]
        rows = n * n if k == _abi.KERNEL_RW_GAUSS_MIX else (1 if k == _abi.KERNEL_MALA else n)
        out = np.empty((rows, self.C))
        self.ck(self.lib.extmcmc_get_eps(self.h, u - 1, _abi.dptr(out)))
        return out

    def eval_grad(self):
        ll, g = np.empty(self.C), np.empty((self.p, self.C))
        self.ck(self.lib.extmcmc_eval_grad(self.h, _abi.dptr(ll), _abi.dptr(g)))
        return ll, g

    def adapt_state(self, u):
        n = self.p_u[u - 1]
        mean, cov = np.empty((n, self.C)), np.empty((n * n, self.C))
        self.ck(self.lib.extmcmc_get_adapt_state(self.h, u - 1, _abi.dptr(mean), _abi.dptr(cov)))
        return mean, cov

    def eval_loglik(self):
        out = np.empty(self.C)
        self.ck(self.lib.extmcmc_eval_loglik(self.h, _abi.dptr(out)))
        return out

    def variant(self):
        return self.lib.extmcmc_sweep_variant_name(self.h).decode()


def compare_histories(o, g, tie_tol=1e-12):
    """o: oracle run dict (with llr, exp_draws); g: GPU run dict.  Returns a report."""
    acc_o, acc_g = o["accepted"].astype(bool), g["accepted"].astype(bool)
    n, Cn = acc_o.shape
    mism = acc_o != acc_g
    # per chain: everything before its first decision mismatch must agree exactly
    first_bad = np.where(mism.any(axis=0), mism.argmax(axis=0), n)
    valid = np.arange(n)[:, None] < first_bad[None, :]          # rows strictly before divergence
    near_tie = 0
    for c in np.nonzero(first_bad < n)[0]:
        s = first_bad[c]
        margin = abs(o["exp_draws"][s, c] + o["llr"][s, c])
        scale = max(abs(o["ll_prop"][s, c]), 1.0)
        if margin <= tie_tol * scale:
            near_tie += 1
    hard = int((first_bad < n).sum()) - near_tie
    vt = np.broadcast_to(valid[:, None, :], o["theta"].shape)
    # (NaN proposals -- a Sigma_B that is not positive definite: PosDefException in the reference --
    # must be NaN on both sides)
    theta_ok = bool(np.array_equal(o["theta"][vt], g["theta"][vt], equal_nan=True) and
                    np.array_equal(o["theta_prop"][vt], g["theta_prop"][vt], equal_nan=True))
    fin = np.isfinite(o["ll_prop"]) & valid
    rel = np.abs(o["ll_prop"][fin] - g["ll_prop"][fin]) / np.maximum(np.abs(o["ll_prop"][fin]), 1e-300)
    fin2 = np.isfinite(o["ll"]) & valid
    rel2 = np.abs(o["ll"][fin2] - g["ll"][fin2]) / np.maximum(np.abs(o["ll"][fin2]), 1e-300)
    return dict(accept_mismatch=hard, near_ties=near_tie, theta_bitexact=theta_ok,
                ll_rel_err=float(max(rel.max() if rel.size else 0.0, rel2.max() if rel2.size else 0.0)),
                chains_diverged=int((first_bad < n).sum()), n_steps=n, n_chains=Cn,
                accept_rate=float(acc_o.mean()))


def replay_compare(x, n_chains, n_iters, seed=1, updates=None, theta_init=None, exclude=(),
                   block=None, check_state=True, law=None, y=None, **cfg_kw):
    """Oracle (own Philox stream, recording) vs GPU (replaying) on the same inputs."""
    law = law if law is not None else em.GsnTargetLaw([0.0], [[1.0]])
    updates = updates if updates is not None else cfg2_updates(eps0=0.05, scale=5e-3, k=10, offset=2.0)
    theta_init = theta_init if theta_init is not None else theta_init_for(x, n_chains)
    steps = list(em.MCMCSchedule(n_iters, len(updates), exclude))
    o = orc.Oracle(law, updates, x, theta_init, n_chains, seed=seed,
                   roll_window=cfg_kw.get("roll_window", 100), y=y)
    ro = o.run(steps, n_threads=8)
    g = GpuSession(law, updates, x, theta_init, n_chains, seed=seed, n_steps_hint=len(steps), y=y, **cfg_kw)
    block = block or len(steps)
    parts = []
    for b in range(0, len(steps), block):
        sl = slice(b, min(b + block, len(steps)))
        parts.append(g.run(steps[sl], replay=(ro["proposals"][sl], ro["exp_draws"][sl])))
    rg = {k: np.concatenate([q[k] for q in parts]) for k in ("theta", "theta_prop", "ll", "ll_prop", "accepted")}
    rep = compare_histories(ro, rg)
    rep["variant"] = g.variant()
    if check_state and rep["chains_diverged"] == 0:
        so, sg = o.stats(), g.stats()
        with_state = [u for u in range(1, len(updates) + 1) if g.kernels[u - 1] != _abi.KERNEL_RW_GAUSS]
        rep["eps_bitexact"] = all(np.array_equal(o.eps(u), g.eps(u)) for u in with_state)
        rep["eps_max_rel"] = max([float(np.max(np.abs(o.eps(u) - g.eps(u)) / np.maximum(np.abs(o.eps(u)), 1e-300)))
                                  for u in with_state] or [0.0])
        rep["mean_bitexact"] = bool(np.array_equal(so["mean"], sg["mean"]))
        cov_o = np.einsum("aac->ac", so["cov"]) if g.stats_mode == 1 else so["cov"]
        rep["cov_bitexact"] = bool(np.array_equal(cov_o, sg["cov"]))
        rep["rolling_ar_bitexact"] = bool(np.array_equal(so["rolling_ar"], sg["rolling_ar"]))
        rep["counts_equal"] = bool(np.array_equal(so["n_accept"], sg["n_accept"]) and
                                   np.array_equal(so["n_prop"], sg["n_prop"]))
        tho, llo = o.state()
        thg, llg = g.state()
        rep["final_state_bitexact"] = bool(np.array_equal(tho, thg))
    g.close()
    return rep
