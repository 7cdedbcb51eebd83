"""Importable alias of the product package.

The package directory is ``stable-diffusion-pytorch_b200/`` (the name the project layout
prescribes); a hyphen is not a legal Python identifier, so this shim makes it importable as
``stable_diffusion_pytorch_b200`` by pointing ``__path__`` at the real directory and executing
its ``__init__``.  No code lives here.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "stable-diffusion-pytorch_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f, _os
