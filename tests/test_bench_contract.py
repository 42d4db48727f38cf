"""bench.py's reference arm (`--impl reference`) runs on host cores only, so its part of the driver contract can be checked
here: one JSON line from rank 0, none from the other ranks, the keys the driver reads, all host threads even under torchrun
(which forces OMP_NUM_THREADS=1)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_under_torchrun_prints_one_line():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29591",
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-scale", "0.001"]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sweeps/s" and d["n_gpus"] == 2 and d["higher_is_better"] is True
    assert d["metric"].startswith("Gibbs sweeps/sec") and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == len(os.sched_getaffinity(0))
