"""bdf_b200 — B200-native Gibbs-sampling engine behind BayesianDataFusion.jl's API (latent-factor hot path).

Host side = a Python mirror of the reference's Julia interface (RelationData / Entity / Relation / macau) over the
C ABI of libbdf_b200.so (include/bdf_b200.h). Importing this package does not need a GPU; creating an Engine does.
"""
from . import _lib  # noqa: F401
from .engine import BDFError, Engine  # noqa: F401

__all__ = ["Engine", "BDFError"]
