"""One GPU playing rank R of an N-GPU run on the C2 workload: times the row kernel of each mode on that rank's shard (no peer stores,
no collectives) — isolates shard-size effects (chunking of heavy rows, wave quantisation) from communication.
    python tools/shard_probe.py WORLD [RANK] [balanced|cyclic] [D]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bdf_b200
from bdf_b200.shard import balanced_partition
from bench import synth, N_USERS, N_ITEMS, NNZ, SEED, ALPHA

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
part = sys.argv[3] if len(sys.argv) > 3 else "balanced"
D = int(sys.argv[4]) if len(sys.argv) > 4 else 100
ids, vals = synth(N_USERS, N_ITEMS, NNZ, SEED)
ids, vals = ids[1_000_000:], vals[1_000_000:]
eng = bdf_b200.Engine(D, rank=rank, world=world)
if part == "balanced" and world > 1:
    e1 = eng.add_entity_partitioned(N_USERS, balanced_partition(np.bincount(ids[:, 0] - 1, minlength=N_USERS), world, 2.0 * D))
    e2 = eng.add_entity_partitioned(N_ITEMS, balanced_partition(np.bincount(ids[:, 1] - 1, minlength=N_ITEMS), world, 2.0 * D))
else:
    e1, e2 = eng.add_entity(N_USERS), eng.add_entity(N_ITEMS)
rel = eng.add_relation([e1, e2], ids, vals)
eng.set_relation_params(rel, ALPHA, float(vals.mean()))
rng = np.random.default_rng(0)
eng.set_factors(e1, rng.standard_normal((N_USERS, D)) * 0.3)
eng.set_factors(e2, rng.standard_normal((N_ITEMS, D)) * 0.3)
for e, name in ((e1, "users"), (e2, "items")):
    ts = []
    for _ in range(4):
        eng.synchronize(); t0 = time.perf_counter(); eng.step_sample(e); eng.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    ph, n = eng.debug_phase_clocks(e)
    print(f"world={world} rank={rank} {part} D={D} {name}: {np.mean(ts[1:]):.3f} ms; items finalising a row {n}; phases {dict((k, int(v)) for k, v in ph.items())}", flush=True)
eng.close()
