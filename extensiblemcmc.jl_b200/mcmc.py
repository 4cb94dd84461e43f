"""MCMC struct and init! (src/mcmc.jl)."""
from .schedule import MCMCSchedule
from .types import MCMCBackend, isdecorator
from . import workspaces as W


def strip_decorators(updt_and_decor):                       # mcmc.jl:56
    return [u for u in updt_and_decor if not isdecorator(u)]


def get_decorators(updt_and_decor):                         # mcmc.jl:63
    return [u for u in updt_and_decor if isdecorator(u)]


class MCMC:
    """MCMC(updates_and_decorators; backend) -- mcmc.jl:32-49."""

    def __init__(self, updt_and_decor, backend=None):
        if backend is None:
            raise TypeError(
                "MCMC(...; backend=CUDAMCMCBackend(...)) is required: the reference's default "
                "GenericMCMCBackend (CPU) is not part of this package")
        assert isinstance(backend, MCMCBackend)
        self.updates = strip_decorators(list(updt_and_decor))
        self.updates_and_decorators = list(updt_and_decor)
        self.backend = backend
        self.schedule = None
        self.workspace = None


def extra_schedule_params(workspace, updates_and_decorators, **kwargs):   # mcmc.jl:111-117
    return {}


def init_(mcmc, num_mcmc_steps, data, theta_init, exclude_updates=(), **kwargs):
    """init!(mcmc, ...) -- mcmc.jl:83-109: global workspace first, then the schedule."""
    mcmc.workspace = W.init_global_workspace(
        mcmc.backend, num_mcmc_steps, mcmc.updates, data, theta_init, **kwargs)
    mcmc.schedule = MCMCSchedule(
        num_mcmc_steps, len(mcmc.updates), exclude_updates,
        **extra_schedule_params(mcmc.workspace, mcmc.updates_and_decorators, **kwargs))
