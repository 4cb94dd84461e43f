"""Import shim: the package directory is named after the reference repository
(`extensiblemcmc.jl_b200/`), which is not a valid Python identifier, so this module
exposes it as the importable package `extensiblemcmc_jl_b200`."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "extensiblemcmc.jl_b200")]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _os, _f
