"""scdrs_b200 -- B200-native implementation of the scDRS per-cell scoring hot path.

Mirrors the reference's public surface for that path (``scdrs/__init__.py:1-6``):
``scdrs_b200.score_cell``, ``scdrs_b200.preprocess``, ``scdrs_b200.method``, ``scdrs_b200.pp``,
``scdrs_b200.util``.  The arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/scdrs_b200.h``; there is no CPU fallback.
"""
from . import method, pp, util  # noqa: F401
from ._device import clear_device_cache, invalidate_uploads, set_default_device, set_spmm_mode  # noqa: F401
from .anndata_lite import AnnData, read_h5ad  # noqa: F401
from .method import score_cell, score_traits  # noqa: F401
from .pp import preprocess  # noqa: F401

__version__ = "0.1.0"
