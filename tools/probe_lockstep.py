"""Lockstep check for the row kernel (needs a -DBDF_DEBUG library for the phase clocks): rows of equal length, 1, 2, 4 and 16 waves of
592 rows (4 per SM) — does a wave whose four co-resident rows are all in the same phase cost more per row than a steady stream? Results in
profiles/README.md ("Lockstep check")."""
import sys, time
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
    ts = []
    for _ in range(5):
        eng.synchronize(); t0 = time.perf_counter(); eng.step_sample(e1); eng.synchronize(); ts.append(time.perf_counter() - t0)
    ph, n = eng.debug_phase_clocks(e1)
    t = np.median(ts)
    print(f"D={D} rows={nrows} ({nrows/592:.1f} waves) obs/row={nobs}: {t*1e3:.3f} ms = {t*1.965e9/(nrows/148):.0f} SM-cycles per row; phases", {k: int(x) for k, x in ph.items()}, flush=True)
    eng.close()
for nr in (592, 1184, 2368, 592 * 16):
    probe(100, nr, 208)
