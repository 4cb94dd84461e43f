"""Random-walk transition kernels (src/transition_kernels/random_walk.jl).

These are host-side parameter holders; sampling and `logpdf(rw, theta, theta°)` run on the
device (csrc/step_kernels.cu) for every chain at once.
"""
import numpy as np

from . import _abi
from .types import TransitionKernel


class RandomWalk(TransitionKernel):
    pass


class UniformRandomWalk(RandomWalk):
    """UniformRandomWalk(eps, pos) -- random_walk.jl:45-53.  theta°_i = theta_i + U_i, or
    theta_i * exp(U_i) for coordinates flagged positive, U_i ~ Unif(-eps_i, eps_i)."""

    abi_kernel = _abi.KERNEL_RW_UNIFORM

    def __init__(self, eps, pos=None):
        self.eps = np.atleast_1d(np.asarray(eps, dtype=np.float64)).copy()
        assert np.all(self.eps > 0.0)                       # random_walk.jl:50
        self.pos = (np.zeros(self.eps.shape, dtype=bool) if pos is None
                    else np.atleast_1d(np.asarray(pos, dtype=bool)).copy())
        assert self.pos.shape == self.eps.shape

    def __len__(self):
        return self.eps.shape[0]

    def abi_step(self):
        return self.eps


class GaussianRandomWalk(RandomWalk):
    """GaussianRandomWalk(Sigma, pos) -- random_walk.jl:123-134."""

    abi_kernel = _abi.KERNEL_RW_GAUSS

    def __init__(self, Sigma, pos=None):
        S = np.atleast_2d(np.asarray(Sigma, dtype=np.float64)).copy()
        assert S.shape[0] == S.shape[1]
        self.Sigma = S
        self.pos = np.zeros(S.shape[0], dtype=bool) if pos is None else np.asarray(pos, dtype=bool).copy()

    def __len__(self):
        return self.Sigma.shape[0]

    def abi_step(self):
        return np.ascontiguousarray(self.Sigma.T).ravel()


class GaussianRandomWalkMix(RandomWalk):
    """GaussianRandomWalkMix(Sigma_A, Sigma_B, lambda, pos) -- random_walk.jl:193-206."""

    abi_kernel = _abi.KERNEL_RW_GAUSS_MIX

    def __init__(self, Sigma_A, Sigma_B, lam=0.5, pos=None):
        assert 0.0 <= lam <= 1.0
        self.gsn_A = GaussianRandomWalk(Sigma_A, pos)
        self.gsn_B = GaussianRandomWalk(Sigma_B, pos)
        assert self.gsn_A.Sigma.shape == self.gsn_B.Sigma.shape
        self.lam = float(lam)
        self.pos = self.gsn_A.pos

    def __len__(self):
        return len(self.gsn_A)

    def abi_step(self):
        return np.concatenate([self.gsn_A.abi_step(), self.gsn_B.abi_step(), [self.lam]])
