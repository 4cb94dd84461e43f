"""Priors (src/priors.jl).  Evaluated on the device on the update's own coordinates only
(src/run.jl:374-385, src/updates.jl:104).  Arbitrary `dist` objects cannot cross the C ABI:
the enumerated families below are supported, anything else raises at workspace creation
(the reference's convention for a missing method is error("... not implemented"))."""
import numpy as np

from . import _abi


class Prior:                                          # priors.jl:4
    def to_abi(self):
        raise NotImplementedError(f"logpdf not implemented for prior {type(self).__name__} on the GPU path")


class ImproperPrior(Prior):                           # priors.jl:18-19
    def to_abi(self):
        return _abi.PRIOR_IMPROPER, np.zeros(0)


class ImproperPosPrior(Prior):                        # priors.jl:25-26
    def to_abi(self):
        return _abi.PRIOR_IMPROPER_POS, np.zeros(0)


class Normal:
    """Stand-in for Distributions.Normal(mu, sigma) inside StandardPrior."""
    def __init__(self, mu=0.0, sigma=1.0):
        self.mu, self.sigma = float(mu), float(sigma)


class Gamma:
    """Stand-in for Distributions.Gamma(shape, scale)."""
    def __init__(self, shape=1.0, scale=1.0):
        self.shape, self.scale = float(shape), float(scale)


class Uniform:
    """Stand-in for Distributions.Uniform(a, b)."""
    def __init__(self, a=0.0, b=1.0):
        self.a, self.b = float(a), float(b)


class Exponential:
    """Stand-in for Distributions.Exponential(scale)."""
    def __init__(self, scale=1.0):
        self.scale = float(scale)


class InverseGamma:
    """Stand-in for Distributions.InverseGamma(shape, scale)."""
    def __init__(self, shape=1.0, scale=1.0):
        self.shape, self.scale = float(shape), float(scale)


class Beta:
    """Stand-in for Distributions.Beta(alpha, beta)."""
    def __init__(self, alpha=1.0, beta=1.0):
        self.alpha, self.beta = float(alpha), float(beta)


class LogNormal:
    """Stand-in for Distributions.LogNormal(mu, sigma)."""
    def __init__(self, mu=0.0, sigma=1.0):
        self.mu, self.sigma = float(mu), float(sigma)


class Cauchy:
    """Stand-in for Distributions.Cauchy(mu, sigma)."""
    def __init__(self, mu=0.0, sigma=1.0):
        self.mu, self.sigma = float(mu), float(sigma)


class MvNormal:
    """Stand-in for Distributions.MvNormal(mu, Sigma): a joint prior over the whole coordinate
    block of an update (priors.jl:35-39 wraps any Distributions object).  The host factorises
    Sigma = L L'; the device whitens with L, as PDMats does."""
    def __init__(self, mu, Sigma):
        self.mu = np.atleast_1d(np.asarray(mu, dtype=np.float64)).copy()
        self.Sigma = np.atleast_2d(np.asarray(Sigma, dtype=np.float64)).copy()
        n = self.mu.size
        if self.Sigma.shape != (n, n):
            raise ValueError("MvNormal: Sigma must be length(mu) x length(mu)")
        self.L = np.linalg.cholesky(self.Sigma)       # raises if not positive definite (PosDefException)


class StandardPrior(Prior):                           # priors.jl:35-39
    """StandardPrior(dist): a univariate `dist` is applied independently to each coordinate of
    the update (an iid product); an MvNormal covers the update's coordinate block jointly."""
    def __init__(self, dist):
        self.dist = dist

    def to_abi(self):
        d = self.dist
        if isinstance(d, Normal):
            return _abi.PRIOR_NORMAL, np.array([d.mu, d.sigma])
        if isinstance(d, Gamma):
            return _abi.PRIOR_GAMMA, np.array([d.shape, d.scale])
        if isinstance(d, Uniform):
            return _abi.PRIOR_UNIFORM, np.array([d.a, d.b])
        if isinstance(d, Exponential):
            return _abi.PRIOR_EXPONENTIAL, np.array([d.scale])
        if isinstance(d, InverseGamma):
            return _abi.PRIOR_INV_GAMMA, np.array([d.shape, d.scale])
        if isinstance(d, Beta):
            return _abi.PRIOR_BETA, np.array([d.alpha, d.beta])
        if isinstance(d, LogNormal):
            return _abi.PRIOR_LOGNORMAL, np.array([d.mu, d.sigma])
        if isinstance(d, Cauchy):
            return _abi.PRIOR_CAUCHY, np.array([d.mu, d.sigma])
        if isinstance(d, MvNormal):
            # {mu[n], L[n*n] column-major}
            return _abi.PRIOR_MVNORMAL, np.concatenate([d.mu, np.ascontiguousarray(d.L.T).ravel()])
        raise NotImplementedError(
            f"StandardPrior({type(d).__name__}) is not implemented on the GPU path (supported: Normal, "
            "Gamma, Uniform, Exponential, InverseGamma, Beta, LogNormal, Cauchy, MvNormal)")


class ProductPrior(Prior):                            # priors.jl:60-88
    """ProductPrior(dists, dims): factor k applies to the next dims[k] coordinates of the update
    (consecutive index groups, priors.jl:66-78).  Each factor is an ImproperPrior /
    ImproperPosPrior or one of the distribution stand-ins above (iid within its group)."""

    def __init__(self, dists, dims):
        self.dists, self.dims = tuple(dists), tuple(int(d) for d in dims)
        assert len(self.dists) == len(self.dims)

    def to_abi(self):
        out = [float(len(self.dists))]
        for dist, dim in zip(self.dists, self.dims):
            pr = dist if isinstance(dist, Prior) else StandardPrior(dist)
            kind, pp = pr.to_abi()
            if kind in (_abi.PRIOR_PRODUCT, _abi.PRIOR_MVNORMAL):
                raise NotImplementedError("ProductPrior factors must be iid families on the GPU path")
            pp = list(pp) + [0.0, 0.0]
            out += [float(kind), float(dim), pp[0], pp[1]]
        return _abi.PRIOR_PRODUCT, np.array(out)
