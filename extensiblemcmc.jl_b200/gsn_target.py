"""GsnTargetLaw -- the reference's shipped example law (src/example/gsn_target.jl:1-29):
a multivariate normal whose mean and covariance are the parameters, theta = [mu; vec(Sigma)].
On the GPU the law is an enumerated model id; `set_parameters!`/`loglikelihood` run on the
device for every chain (csrc/sweep_gsn1d.cu)."""
import numpy as np

from . import _abi


class GsnTargetLaw:
    def __init__(self, mu, Sigma=None):
        mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        d = mu.shape[0]
        Sigma = np.eye(d) if Sigma is None else np.atleast_2d(np.asarray(Sigma, dtype=np.float64))
        assert Sigma.shape == (d, d)
        self.d = d
        self.theta = np.concatenate([mu, Sigma.T.ravel()])       # gsn_target.jl:6-9 (vec = column-major)

    def abi_law(self):
        if self.d == 1:
            return _abi.LAW_GSN_IID_1D
        return _abi.LAW_GSN_MV

    @property
    def obs_dim(self):
        return self.d

    @property
    def n_params(self):
        return self.d * (self.d + 1)
