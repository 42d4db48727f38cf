import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import bdf_b200
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError: pass
def peek(tag):
    e = rt.cudaPeekAtLastError() if rt else -1
    print(tag, "->", e, flush=True)
rows = np.concatenate([np.arange(1, 201), np.arange(151, 351)]).astype(np.int32)
cols = np.concatenate([np.arange(151, 351), np.arange(1, 400, 2)]).astype(np.int32)
eng = bdf_b200.Engine(8); peek("create")
e1, e2 = eng.add_entity(350), eng.add_entity(7); peek("add_entity")
rng = np.random.default_rng(0)
ids = np.stack([rng.integers(1, 351, 1000), rng.integers(1, 8, 1000)], 1)
rel = eng.add_relation([e1, e2], ids, rng.standard_normal(1000)); peek("add_relation")
eng.set_features(e1, bdf_b200.SparseBinMatrix(rows, cols, 350, 399)); peek("set_features")
p, i = eng.debug_features_csr(e1, False, 400); peek("debug_csr")
y = eng.spmm(e1, rng.standard_normal((399, 3))); peek("spmm")
eng.close(); peek("close")
