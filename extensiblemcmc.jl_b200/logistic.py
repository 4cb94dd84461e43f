"""LogisticLaw -- Bayesian logistic regression of BASELINE cfg 3 (no counterpart in the
reference; build-defined): theta = beta[d], y_i ~ Bernoulli(sigmoid(x_i . beta)).
Data: dict(P=LogisticLaw(d), obs=X [N, d], y=y [N]).  Log-likelihood and gradient run as FP64
tensor-core GEMMs on the device (csrc/sweep_logistic.cu)."""
import numpy as np

from . import _abi


class LogisticLaw:
    def __init__(self, d):
        self.d = int(d)
        assert 1 <= self.d

    def abi_law(self):
        return _abi.LAW_LOGISTIC

    @property
    def obs_dim(self):
        return self.d

    @property
    def n_params(self):
        return self.d

    def abi_y(self, data):
        y = data["y"] if isinstance(data, dict) else data.y
        return np.ascontiguousarray(np.asarray(y, dtype=np.float64))
