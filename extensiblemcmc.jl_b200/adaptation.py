"""Adaptation schemes (src/transition_kernels/adaptation.jl).

`AdaptationUnifRW` keeps the reference's constructor semantics (defaults, scalar/vector
promotion, `==`, `isequal_except` -- test/runtests.jl:34-85).  The adaptation itself
(register!/time_to_update/readjust!, adaptation.jl:273-329) runs per chain on the device.
"""
import numpy as np

from . import _abi
from .types import Adaptation


class NoAdaptation(Adaptation):                      # adaptation.jl:26
    def to_abi(self):
        return _abi.Adapt(_abi.ADAPT_NONE, 100, 0.234, 1.0, 1e-12, 1e7, 1e2)


def _is_scalar(x):
    return np.ndim(x) == 0


class AdaptationUnifRW(Adaptation):
    """AdaptationUnifRW(theta; adapt_every_k_steps=100, target_accpt_rate=0.234, scale=1.0,
    min=1e-12, max=1e7, offset=1e2) -- adaptation.jl:51-199.

    `theta` only fixes the number of coordinates N and whether scale/min/max/offset are
    promoted to length-N vectors: they stay scalars unless at least one of them is given
    as a vector (adaptation.jl:94-104)."""

    _fields = ("proposed", "accepted", "target_accpt_rate", "adapt_every_k_steps",
               "scale", "min", "max", "offset", "N")

    def __init__(self, theta, **kwargs):
        allowed = {"adapt_every_k_steps", "target_accpt_rate", "scale", "min", "max", "offset"}
        unknown = set(kwargs) - allowed
        if unknown:
            raise TypeError(f"unknown keyword(s) {sorted(unknown)}")
        self.proposed = 0
        self.accepted = 0
        self.target_accpt_rate = float(np.asarray(kwargs.get("target_accpt_rate", 0.234)).reshape(-1)[0])
        self.adapt_every_k_steps = int(np.asarray(kwargs.get("adapt_every_k_steps", 100)).reshape(-1)[0])
        self.N = int(np.size(theta))
        vec = {k: kwargs.get(k, d) for k, d in
               (("scale", 1.0), ("min", 1e-12), ("max", 1e7), ("offset", 1e2))}
        given = [kwargs[k] for k in ("scale", "min", "max", "offset") if k in kwargs]
        lengths = {int(np.size(v)) for v in given}
        assert len(lengths) <= 2                         # adaptation.jl:96-97
        nonscalar = bool(lengths) and max(lengths) > 1
        for k, v in vec.items():
            if nonscalar:
                a = np.asarray(v, dtype=np.float64).reshape(-1)
                a = np.full(self.N, a[0]) if a.size == 1 else a.copy()
                assert a.size == self.N
                setattr(self, k, a)
            else:
                setattr(self, k, float(np.asarray(v).reshape(-1)[0]))

    def __eq__(self, other):                             # adaptation.jl:206-213
        return isinstance(other, AdaptationUnifRW) and isequal_except(self, other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def to_abi(self):
        # Only scalar scale/min/max/offset work on the reference's run path (compute_delta
        # does max(1.0, vector), a MethodError -- adaptation.jl:312-319); same here.
        for k in ("scale", "min", "max", "offset"):
            if not _is_scalar(getattr(self, k)):
                raise NotImplementedError(
                    f"AdaptationUnifRW.{k} given as a vector: the reference's readjust! only "
                    "works for scalar parameters; not implemented on the GPU path")
        return _abi.Adapt(_abi.ADAPT_UNIF_RW, self.adapt_every_k_steps, self.target_accpt_rate,
                          self.scale, self.min, self.max, self.offset)


def isequal_except(a, b, *args):                         # adaptation.jl:223-235
    if type(a) is not type(b):
        return False
    for fn in AdaptationUnifRW._fields:
        if fn in args:
            continue
        x, y = getattr(a, fn), getattr(b, fn)
        if _is_scalar(x) != _is_scalar(y):               # T != S in the reference
            return False
        if not np.array_equal(np.asarray(x), np.asarray(y)):
            return False
    return True


class HaarioTypeAdaptation(Adaptation):
    """HaarioTypeAdaptation(state; adapt_every_k_steps=100, scale=2.38^2, f) --
    adaptation.jl:372-397.  Runs on the device (running mean / covariance of the transformed
    sub-state, Sigma_B readjusted every k own turns); the weight schedule `f(lambda, N, iter)` is a
    host closure: lambda is shared by all chains and N, iter are schedule facts, so the library calls
    `f` back on the host at every readjustment (extmcmc_set_lambda_fn)."""

    def __init__(self, state, adapt_every_k_steps=100, scale=2.38 ** 2, f=None):
        n = int(np.size(state))
        self.mean = np.zeros(n)
        self.cov = np.zeros((n, n))
        self.adapt_every_k_steps = int(adapt_every_k_steps)
        self.scale = float(scale)
        self.N, self.M = 1, 0
        self._default_f = f is None
        self.f = f if f is not None else (lambda x, y, z: x)

    def lambda_callback(self):
        """ctypes callback for extmcmc_set_lambda_fn (None for the default, constant lambda)."""
        if self._default_f:
            return None
        f = self.f
        return _abi.LambdaFn(lambda lam, N, it, _user: float(f(lam, int(N), int(it))))

    def to_abi(self):
        # NB the reference's readjust! ignores `scale` and uses 2.38^2 / length(rw) (adaptation.jl:423)
        return _abi.Adapt(_abi.ADAPT_HAARIO, self.adapt_every_k_steps, 0.0, self.scale, 0.0, 0.0, 0.0)


class AdaptationMALA(AdaptationUnifRW):
    """Acceptance-rate targeting of the MALA step size tau with the reference's +-delta rule
    (compute_delta / compute_eps, adaptation.jl:312-329); default target 0.574.  The reference
    only has a TODO for adaptive MALA (adaptation.jl:8-10)."""

    def __init__(self, **kwargs):
        kwargs.setdefault("target_accpt_rate", 0.574)
        super().__init__([0.0], **kwargs)

    def to_abi(self):
        a = super().to_abi()
        a.kind = _abi.ADAPT_MALA
        return a
