"""(FᵀF + λI)·X on the C3 feature matrix for full-width and windowed operands (one GPU): milliseconds per application."""
import sys

import numpy as np

sys.path.insert(0, ".")
import bdf_b200
from tools.workloads import c3_macau

w = c3_macau(1.0, 32)
eng = bdf_b200.Engine(32)
e1, e2 = eng.add_entity(w["N"]), eng.add_entity(w["NT"])
rel = eng.add_relation([e1, e2], w["ids"], w["vals"])
eng.set_features(e1, bdf_b200.SparseBinMatrix(w["rows"], w["cols"], w["N"], w["NUMF"]))
eng.set_beta(e1, np.random.default_rng(0).standard_normal((w["NUMF"], 32)))
print(f"in-order kernel pair (bdf_ata_mul, F*beta), 32 columns: {eng.debug_ata_time(e1, 20, -1):.4f} ms per application", flush=True)
for nc in (0, 16, 8, 4):
    print(f"order-free kernel pair (CG iteration), columns {nc or 32}: {eng.debug_ata_time(e1, 20, nc):.4f} ms per application", flush=True)
eng.close()
