"""pytorch-vae_b200: B200-native (sm_100a) vector-quantizer hot path of jluuser/PyTorch-VAE.

Drop-in for ``models/vq_vae.py``'s ``VectorQuantizerEMA``; see INTEGRATION.md.
Importing this package loads ``lib/libvqb200.so`` and raises if it is missing --
there is no fallback implementation.
"""
from . import _cabi, ops  # noqa: F401  (loads the shared library; ImportError if absent)
from .quantizer import VectorQuantizer, VectorQuantizerEMA  # noqa: F401
from .integration import install, uninstall  # noqa: F401
from .graphs import GraphedForward, GraphedTrainStep  # noqa: F401
from .kmeans import kmeans_fit, rvq_kmeans_fit  # noqa: F401

__all__ = ["VectorQuantizerEMA", "VectorQuantizer", "GraphedForward", "GraphedTrainStep", "kmeans_fit", "rvq_kmeans_fit", "install",
           "uninstall", "ops"]
