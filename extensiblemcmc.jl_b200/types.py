"""Abstract type tree of the reference (src/types.jl:8-125), kept so that user code and
the backend seam read the same: `MCMC(updates; backend=CUDAMCMCBackend(...))`."""


class MCMCUpdate:                                  # types.jl:8
    pass


class MCMCParamUpdate(MCMCUpdate):                 # types.jl:16
    pass


class MCMCGradientBasedUpdate(MCMCParamUpdate):    # types.jl:24
    pass


class MCMCConjugateParamUpdate(MCMCParamUpdate):   # types.jl:32
    pass


class MCMCImputation(MCMCUpdate):                  # types.jl:49
    pass


class MCMCUpdateDecorator:                         # types.jl:58
    pass


def isdecorator(u):                                # types.jl:63-65
    return isinstance(u, MCMCUpdateDecorator)


class Workspace:                                   # types.jl:71
    pass


class GlobalWorkspace(Workspace):                  # types.jl:79
    pass


class LocalWorkspace(Workspace):                   # types.jl:87
    pass


class TransitionKernel:                            # types.jl:95
    pass


class Adaptation:                                  # types.jl:104
    pass


class MCMCBackend:                                 # types.jl:110
    """Plugin seam of the reference: `init_global_workspace(::Backend, ...)` and
    `create_workspace(::Backend, ...)` dispatch on it (src/workspaces.jl:38-47,280-287)."""


class GenericMCMCBackend(MCMCBackend):             # types.jl:117
    """The reference's CPU backend.  It is NOT provided here: this repository ships only
    the B200 path (no CPU fallback); selecting it raises in `init_global_workspace`."""


class ChainStats:                                  # types.jl:125
    pass


class Previous:                                    # flags, types.jl:128-142
    pass


class Proposal:
    pass


class PreMCMCStep:
    pass


class PostMCMCStep:
    pass


PREVIOUS, PROPOSAL, PRESTEP, POSTSTEP = Previous(), Proposal(), PreMCMCStep(), PostMCMCStep()
