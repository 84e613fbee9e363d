"""Import shim: the product package lives in ``diffusion-net_b200/`` (the name the
build contract mandates, which is not a Python identifier).  ``import
diffusion_net_b200`` resolves here and re-exports that directory as this package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "diffusion-net_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _fh
