"""Scratch timing of device-resident sweeps on synthetic BPMF data (not the judged bench; see bench.py)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bdf_b200


def synth(n1, n2, nnz, seed, skew=2.5):
    rng = np.random.default_rng(seed)
    i1 = np.floor(n1 * rng.random(nnz) ** skew).astype(np.int64) + 1
    i2 = np.floor(n2 * rng.random(nnz) ** skew).astype(np.int64) + 1
    v = rng.standard_normal(nnz)
    return np.stack([i1, i2], axis=1), v


def run(n1, n2, nnz, D, sweeps=3, skew=2.5):
    ids, v = synth(n1, n2, nnz, 1, skew)
    t0 = time.time()
    eng = bdf_b200.Engine(D)
    e1, e2 = eng.add_entity(n1), eng.add_entity(n2)
    rel = eng.add_relation([e1, e2], ids, v)
    eng.set_relation_params(rel, 1.5, float(v.mean()))
    eng.synchronize()
    t1 = time.time()
    eng.sweep(2)
    eng.synchronize()
    ts = []
    for e in (e1, e2):
        for name, fn in (("sample", eng.step_sample), ("stats", eng.step_nw_stats), ("draw", eng.step_nw_draw)):
            eng.synchronize()
            a = time.time()
            fn(e)
            eng.synchronize()
            ts.append((e, name, (time.time() - a) * 1e3))
    a = time.time()
    eng.sweep(sweeps)
    eng.synchronize()
    dt = (time.time() - a) / sweeps
    flops = 2 * nnz * (D * (D + 1) + 2 * D) + (n1 + n2) * (D**3 / 3 + 2 * D * D)
    print(f"D={D} n=({n1},{n2}) nnz={nnz} skew={skew}: ingest {t1-t0:.2f}s, sweep {dt*1e3:.2f} ms, {1/dt:.2f} sweeps/s, "
          f"{flops/dt/1e12:.2f} TFLOP/s alg; parts(ms): " + ", ".join(f"e{e}.{n}={t:.2f}" for e, n, t in ts), flush=True)
    for e in (e1, e2):
        print('   phase clocks', e, {k: int(v) for k, v in eng.debug_phase_clocks(e)[0].items()}, flush=True)
    eng.close()


if __name__ == "__main__":
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
    for D in (32, 100):
        run(int(480000 * scale), 17800, int(1e8 * scale), D)
        run(int(480000 * scale), 17800, int(1e8 * scale), D, skew=1.0)
