"""Importable alias of the ``pytorch-vae_b200/`` package directory (a hyphen is not a valid
identifier): ``import pytorch_vae_b200`` executes ``pytorch-vae_b200/__init__.py`` with this
module's ``__path__`` pointing there, so submodules resolve to the real files."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pytorch-vae_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
