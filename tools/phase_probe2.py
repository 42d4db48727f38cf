"""Main-loop experiments: phase clocks of the row kernel with the gather or the DMMA accumulate switched off
(BDF_DEBUG_FLAGS: 1 = stop after the syrk, 2 = no DMMA, 4 = no gather). Results of these runs are invalid by design."""
import os
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
    for name, fl in (("full", 0), ("syrk only", 1), ("gather only", 3), ("dmma only", 5)):
        os.environ["BDF_DEBUG_FLAGS"] = str(fl)
        for _ in range(2):
            ph, n = eng.debug_phase_clocks(e1)
        print(f"D={D} rows={nrows} obs/row={nobs} [{name}]: syrk {int(ph['syrk'])} cycles = {ph['syrk'] / (nobs / 4):.0f} per k4-step; total {int(ph['total'])}", flush=True)
    os.environ["BDF_DEBUG_FLAGS"] = "0"
    eng.close()


for D in (int(a) for a in (sys.argv[1:] or ["100", "32"])):
    probe(D, 148, 2080)
    probe(D, 148 * 16, 208)
    probe(D, 148 * 16, 2080)
