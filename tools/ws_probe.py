"""Per-role busy / wait cycles of the persistent warp-specialised row kernel (needs a -DBDF_DEBUG library: BDF_LIB_TAG=dbg
BDF_EXTRA_NVCC=-DBDF_DEBUG python bayesiandatafusion.jl_b200/build.py; run with BDF_B200_LIB=.../libbdf_dbg.so BDF_DEBUG_WS=1)."""
import sys

sys.path.insert(0, ".")
import bdf_b200
from tools.quick_bench import synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
D = 100
n1, n2, nnz = int(480000 * scale), 17800, int(1e8 * scale)
ids, v = synth(n1, n2, nnz, 1)
eng = bdf_b200.Engine(D)
e1, e2 = eng.add_entity(n1), eng.add_entity(n2)
rel = eng.add_relation([e1, e2], ids, v)
eng.set_relation_params(rel, 1.5, float(v.mean()))
eng.sweep(2)
eng.synchronize()
for e, name in ((e1, "users"), (e2, "items")):
    print(f"--- {name} launch", file=sys.stderr, flush=True)
    eng.debug_phase_clocks(e)
eng.close()
