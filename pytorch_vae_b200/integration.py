"""Hooking the B200 quantizer into the reference code base without editing it.

``VQVAE.__init__`` looks ``VectorQuantizerEMA`` up as a module global
(models/vq_vae.py:506), so rebinding that one name makes ``run.py``,
``experiment.py``, ``scripts/extract_code_indices.py`` and
``scripts/decode_with_vqvae.py`` construct this package's quantizer.
"""
from __future__ import annotations

import importlib

from .quantizer import VectorQuantizerEMA

_saved = {}


def install(module="models.vq_vae"):
    """Rebind ``<module>.VectorQuantizerEMA`` to the B200 class. Returns the patched module."""
    mod = importlib.import_module(module) if isinstance(module, str) else module
    if mod.__name__ not in _saved:
        _saved[mod.__name__] = (mod, getattr(mod, "VectorQuantizerEMA", None))
    mod.VectorQuantizerEMA = VectorQuantizerEMA
    return mod


def uninstall(module="models.vq_vae"):
    name = module if isinstance(module, str) else module.__name__
    if name in _saved:
        mod, orig = _saved.pop(name)
        if orig is not None:
            mod.VectorQuantizerEMA = orig
