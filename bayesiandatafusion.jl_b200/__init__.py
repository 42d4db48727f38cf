"""bdf_b200 — B200-native Gibbs-sampling engine behind BayesianDataFusion.jl's API (latent-factor hot path).

Host side = a Python mirror of the reference's Julia interface (RelationData / Entity / Relation / macau) over the
C ABI of libbdf_b200.so (include/bdf_b200.h). Importing this package does not need a GPU; creating an Engine does.
"""
from . import _lib  # noqa: F401
from .engine import BDFError, Engine  # noqa: F401
from . import data_reading  # noqa: F401
from .data_reading import read_binary_float32, read_sparse_binary_matrix, write_binary_matrix  # noqa: F401
from .macau import AUC_ROC, macau  # noqa: F401
from .relation_data import (Entity, IndexedDF, Relation, RelationData, SparseBinMatrix, assignToTest,  # noqa: F401
                            setPrecision, setTest)

__all__ = ["Engine", "BDFError", "macau", "RelationData", "Relation", "Entity", "IndexedDF", "assignToTest", "setTest",
           "setPrecision", "AUC_ROC", "SparseBinMatrix"]
