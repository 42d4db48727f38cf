"""Set up a synthetic BPMF problem and run a few device-resident half-sweeps (target for ncu)."""
import sys

sys.path.insert(0, ".")
import bdf_b200
from tools.quick_bench import synth

D = int(sys.argv[1]); n1 = int(sys.argv[2]); n2 = int(sys.argv[3]); nnz = int(sys.argv[4])
skew = float(sys.argv[5]) if len(sys.argv) > 5 else 2.5
ids, v = synth(n1, n2, nnz, 1, skew)
eng = bdf_b200.Engine(D)
e1, e2 = eng.add_entity(n1), eng.add_entity(n2)
rel = eng.add_relation([e1, e2], ids, v)
eng.set_relation_params(rel, 1.5, float(v.mean()))
eng.sweep(2)
eng.synchronize()
print("done", eng.launches)
