"""Isolated phase latencies: a grid small enough that CTAs do not share SMs."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bdf_b200

def probe(D, nrows, nobs, n2=17800):
    rng = np.random.default_rng(0)
    i1 = np.repeat(np.arange(1, nrows + 1), nobs)
    i2 = rng.integers(1, n2 + 1, nrows * nobs)
    v = rng.standard_normal(nrows * nobs)
    eng = bdf_b200.Engine(D)
    e1, e2 = eng.add_entity(nrows), eng.add_entity(n2)
    rel = eng.add_relation([e1, e2], np.stack([i1, i2], 1), v)
    eng.set_relation_params(rel, 1.5, 0.0)
    eng.set_factors(e2, rng.standard_normal((n2, D)) * 0.3)
    eng.step_sample(e1); eng.synchronize()
    for _ in range(2):
        ph, n = eng.debug_phase_clocks(e1)
    print(f"D={D} rows={nrows} obs/row={nobs}:", {k: int(x) for k, x in ph.items()}, flush=True)
    eng.close()

for D in (100, 32):
    probe(D, 148, 208)
    probe(D, 296, 208)
    probe(D, 148, 2080)
    probe(D, 148 * 16, 208)
