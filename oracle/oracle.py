"""ctypes wrapper of the CPU oracle (oracle/extmcmc_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product never imports this module.
It reuses the ABI struct definitions of the product so that one configuration can be
handed to both sides.
"""
import ctypes as C
import os
import subprocess

import numpy as np

import extensiblemcmc_jl_b200  # noqa: F401  (registers the package)
from extensiblemcmc_jl_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_extmcmc.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    dp, u8p, i64p = _abi.c_double_p, _abi.c_uint8_p, _abi.c_int64_p
    lib.oracle_create.restype = C.c_int32
    lib.oracle_create.argtypes = [C.POINTER(_abi.Config), C.POINTER(_abi.Update), dp, C.c_int64, dp, dp, C.POINTER(H)]
    lib.oracle_destroy.restype = None
    lib.oracle_destroy.argtypes = [H]
    lib.oracle_last_error.restype = C.c_char_p
    lib.oracle_run_block.restype = C.c_int32
    lib.oracle_run_block.argtypes = [H, C.POINTER(_abi.Step), C.c_int32, C.c_int32, C.c_int32,
                                     dp, dp, dp, dp, dp, dp, u8p, dp, C.c_int32]
    lib.oracle_set_lambda_fn.restype = C.c_int32
    lib.oracle_set_lambda_fn.argtypes = [H, C.c_int32, _abi.LambdaFn, C.c_void_p]
    lib.oracle_get_state.restype = C.c_int32
    lib.oracle_get_state.argtypes = [H, dp, dp]
    lib.oracle_get_stats.restype = C.c_int32
    lib.oracle_get_stats.argtypes = [H, dp, dp, dp, i64p, i64p]
    lib.oracle_get_eps.restype = C.c_int32
    lib.oracle_get_eps.argtypes = [H, C.c_int32, dp]
    lib.oracle_get_adapt_state.restype = C.c_int32
    lib.oracle_get_adapt_state.argtypes = [H, C.c_int32, dp, dp]
    lib.oracle_loglik.restype = C.c_int32
    lib.oracle_loglik.argtypes = [H, dp, C.c_int64, dp, C.c_int32]
    lib.oracle_loglik_grad.restype = C.c_int32
    lib.oracle_loglik_grad.argtypes = [H, dp, C.c_int64, dp, dp]
    lib.oracle_philox4x32_10.restype = None
    lib.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.oracle_uniform.restype = C.c_double
    lib.oracle_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int32, C.c_uint32]
    _lib = lib
    return lib


def steps_array(steps):
    """steps: iterable of schedule.Step (1-based) -> (_abi.Step * n)"""
    steps = list(steps)
    arr = (_abi.Step * len(steps))()
    for k, s in enumerate(steps):
        arr[k].mcmciter = s.mcmciter
        arr[k].prev_mcmciter = s.prev_mcmciter if s.prev_mcmciter is not None else 0
        arr[k].pidx = s.pidx - 1
        arr[k].prev_pidx = (s.prev_pidx - 1) if s.prev_pidx is not None else -1
    return arr


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


class Oracle:
    """CPU restatement of the reference sampler for C independent chains."""

    def __init__(self, law, updates, obs, theta_init, n_chains, seed=0, chain_offset=0,
                 roll_window=100, y=None):
        lib = load()
        self.lib = lib
        self.C, self.p, self.NU = int(n_chains), law.n_params, len(updates)
        cfg = _abi.Config()
        cfg.abi_version = _abi.ABI_VERSION
        cfg.n_chains, cfg.chain_offset = self.C, chain_offset
        cfg.n_params, cfg.n_updates = self.p, self.NU
        cfg.law, cfg.obs_dim = law.abi_law(), law.obs_dim
        cfg.seed = seed
        cfg.roll_window = roll_window
        cfg.history_window = 1
        self.cfg = cfg
        arr = (_abi.Update * self.NU)()
        self._keep = []
        self.p_u = []
        for i, u in enumerate(updates):
            a, keep = u.to_abi(self.p)
            arr[i] = a
            self._keep.append(keep)
            self.p_u.append(len(u.coords))
            self.kernels = getattr(self, "kernels", []) + [a.kernel]
        self.p_u_max = max(self.p_u)
        obs = np.ascontiguousarray(np.asarray(obs, dtype=np.float64))
        if obs.ndim == 1:
            obs = obs[:, None]
        theta_init = np.asarray(theta_init, dtype=np.float64)
        if theta_init.ndim == 1:
            theta_init = np.repeat(theta_init[:, None], self.C, axis=1)
        theta_init = np.ascontiguousarray(theta_init)
        assert theta_init.shape == (self.p, self.C)
        self.h = C.c_void_p()
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        rc = lib.oracle_create(C.byref(cfg), arr, _abi.dptr(obs), obs.shape[0], _abi.dptr(y),
                               _abi.dptr(theta_init), C.byref(self.h))
        if rc != 0:
            raise OracleError(rc, lib.oracle_last_error().decode())
        for i, u in enumerate(updates):                 # HaarioTypeAdaptation(...; f = ...)
            cb = getattr(getattr(u, "adpt", None), "lambda_callback", lambda: None)()
            if cb is not None:
                self._keep.append(cb)
                lib.oracle_set_lambda_fn(self.h, i, cb, None)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_destroy(self.h)
            self.h = None

    def run(self, steps, replay=None, n_threads=1, record=True):
        """Run schedule elements; returns dict of histories (+ recorded randomness).
        replay = (proposals [n][p_u_max][C], exp_draws [n][C]) switches to replay mode."""
        steps = list(steps)
        n, Cn, p = len(steps), self.C, self.p
        arr = steps_array(steps)
        out = dict(theta=np.empty((n, p, Cn)), theta_prop=np.empty((n, p, Cn)), ll=np.empty((n, Cn)),
                   ll_prop=np.empty((n, Cn)), accepted=np.empty((n, Cn), dtype=np.uint8),
                   llr=np.empty((n, Cn)))
        if replay is not None:
            props = np.ascontiguousarray(replay[0], dtype=np.float64)
            exps = np.ascontiguousarray(replay[1], dtype=np.float64)
            assert props.shape == (n, self.p_u_max, Cn) and exps.shape == (n, Cn)
            mode = _abi.RNG_REPLAY
        else:
            props = np.full((n, self.p_u_max, Cn), np.nan) if record else None
            exps = np.full((n, Cn), np.nan) if record else None
            mode = _abi.RNG_PHILOX
        rc = self.lib.oracle_run_block(
            self.h, arr, n, mode, self.p_u_max, _abi.dptr(props), _abi.dptr(exps),
            _abi.dptr(out["theta"]), _abi.dptr(out["theta_prop"]), _abi.dptr(out["ll"]),
            _abi.dptr(out["ll_prop"]), out["accepted"].ctypes.data_as(_abi.c_uint8_p),
            _abi.dptr(out["llr"]), n_threads)
        out["rc"] = rc
        if rc not in (0, _abi.EDOMAIN):
            raise OracleError(rc, self.lib.oracle_last_error().decode())
        out["proposals"], out["exp_draws"] = props, exps
        return out

    def state(self):
        th, ll = np.empty((self.p, self.C)), np.empty(self.C)
        self.lib.oracle_get_state(self.h, _abi.dptr(th), _abi.dptr(ll))
        return th, ll

    def stats(self):
        p, Cn, NU = self.p, self.C, self.NU
        mean, cov = np.empty((p, Cn)), np.empty((p * p, Cn))
        ra = np.empty((NU, Cn))
        na, npr = np.empty((NU, Cn), dtype=np.int64), np.empty((NU, Cn), dtype=np.int64)
        self.lib.oracle_get_stats(self.h, _abi.dptr(mean), _abi.dptr(cov), _abi.dptr(ra),
                                  na.ctypes.data_as(_abi.c_int64_p), npr.ctypes.data_as(_abi.c_int64_p))
        return dict(mean=mean, cov=cov.reshape(p, p, Cn).transpose(1, 0, 2), rolling_ar=ra,
                    n_accept=na, n_prop=npr)

    def eps(self, u):
        """eps [p_u, C] (uniform walk) or Sigma_B [p_u^2, C] (Gaussian mixture walk)."""
        n, k = self.p_u[u - 1], self.kernels[u - 1]
        ln = (n if k == _abi.KERNEL_RW_UNIFORM else 1 if k == _abi.KERNEL_MALA
              else (n * n if k == _abi.KERNEL_RW_GAUSS else 2 * n * n + 1))
        out = np.empty((ln, self.C))
        self.lib.oracle_get_eps(self.h, u - 1, _abi.dptr(out))
        return out[n * n:2 * n * n] if k == _abi.KERNEL_RW_GAUSS_MIX else out

    def adapt_state(self, u):
        n = self.p_u[u - 1]
        mean, cov = np.empty((n, self.C)), np.empty((n * n, self.C))
        rc = self.lib.oracle_get_adapt_state(self.h, u - 1, _abi.dptr(mean), _abi.dptr(cov))
        assert rc == 0
        return mean, cov

    def loglik(self, theta, n_threads=1):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        n = theta.shape[1]
        out = np.empty(n)
        self.lib.oracle_loglik(self.h, _abi.dptr(theta), n, _abi.dptr(out), n_threads)
        return out


def _loglik_grad(self, theta):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n = theta.shape[1]
    ll, g = np.empty(n), np.empty((self.p, n))
    rc = self.lib.oracle_loglik_grad(self.h, _abi.dptr(theta), n, _abi.dptr(ll), _abi.dptr(g))
    assert rc == 0
    return ll, g


Oracle.loglik_grad = _loglik_grad


def philox(ctr, key):
    lib = load()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.oracle_philox4x32_10(c, k, o)
    return tuple(o)


def uniform(seed, chain, mcmciter, pidx0, j):
    return load().oracle_uniform(seed, chain, mcmciter, pidx0, j)
