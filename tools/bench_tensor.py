"""C4 of BASELINE.json: 3-mode tensor factorisation, synthetic 20k × 5k × 200 with 50M observed entries, D=30 — the Khatri-Rao
gather path of the row kernel (two partner rows per observation, multiplied while the DMMA fragments are formed); the 200-row
mode has 250k observations per row, so every one of its rows is split over ~31 CTAs. Prints one JSON line: device-resident
sweeps/s, per-mode kernel times and the row kernel's algorithmic FP64 rate and gather bandwidth (SURVEY §8d formulas, K=3)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bdf_b200

from tools.workloads import c4_tensor

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
D = 30
w = c4_tensor(scale)
dims, NNZ, ids, vals = w["dims"], w["NNZ"], w["ids"], w["vals"]
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
eng = bdf_b200.Engine(D)
eng.set_stream(stream.cuda_stream)
ents = [eng.add_entity(d) for d in dims]
rel = eng.add_relation(ents, ids, vals)
eng.set_relation_params(rel, 1.5, float(vals.mean()))
eng.sweep(3)
torch.cuda.synchronize()
K = 10
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); eng.sweep(K); ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / K
kt = []
for e in ents:
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.step_sample(e); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    kt.append(float(np.mean(ts[1:])))
flops = sum(NNZ * (D * (D + 1) + 2 * D + D) + d * (D ** 3 / 3 + 2 * D * D) for d in dims)
gbytes = sum(NNZ * 2 * D * 8 + NNZ * (8 + 8) + 8 * (d + 1) + d * D * 8 for d in dims)
try:
    peak_bw = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    peak_bw = 6650.0
try:
    peak_fl = json.load(open("profiles/fp64_peak_r01.json"))["dmma_sustained_tflops"]
except Exception:
    peak_fl = 37.07
tk = sum(kt) / 1e3
print(json.dumps({"config": f"C4 tensor {dims[0]}x{dims[1]}x{dims[2]}, {NNZ} entries, D={D}", "sweeps_per_s": 1e3 / ms, "ms_per_sweep": ms,
                  "row_kernel_ms_per_mode": kt, "algorithmic_tflops": flops / tk / 1e12, "frac_of_fp64_dmma_peak": flops / tk / 1e12 / peak_fl,
                  "algorithmic_gather_gbs": gbytes / tk / 1e9, "frac_of_hbm_copy_peak": gbytes / tk / 1e9 / peak_bw,
                  "note": "partner matrices (4.8 MB + 1.2 MB + 48 KB) are L2-resident; the compulsory HBM stream is the 16 B/observation CSR payload"}))
eng.close()
