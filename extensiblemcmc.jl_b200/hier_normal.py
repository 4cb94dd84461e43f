"""HierNormalLaw -- the hierarchical normal model of BASELINE cfg 4 (no counterpart in the
reference, whose only shipped law is GsnTargetLaw; build-defined):

    theta = [theta_1, ..., theta_G, mu, tau],   y_gj ~ N(theta_g, 1),   theta_g ~ N(mu, tau^2).

The hierarchical term p(theta_g | mu, tau) is part of the law's log-likelihood because the
reference's priors only ever see the coordinates of their own update (src/run.jl:374-385).
Data: dict(P=HierNormalLaw(G), obs=y, groups=g) with g the 0-based group of every observation
(sorted ascending)."""
import numpy as np

from . import _abi


class HierNormalLaw:
    def __init__(self, n_groups):
        self.G = int(n_groups)
        assert self.G >= 1

    def abi_law(self):
        return _abi.LAW_HIER_NORMAL

    @property
    def obs_dim(self):
        return 1

    @property
    def n_params(self):
        return self.G + 2

    def abi_y(self, data):
        g = data["groups"] if isinstance(data, dict) else data.groups
        g = np.ascontiguousarray(np.asarray(g, dtype=np.float64))
        if np.any(np.diff(g) < 0):
            raise ValueError("observations must be sorted by group")
        return g
