"""Multi-GPU parity on real devices (skipped on a box with fewer than two GPUs): tools/mgpu_check.py under torchrun — the N-GPU
sharded sweep with the fused peer-store all-gather (drawn rows written straight into every peer replica over NVLink by the row
kernel) must equal the same sweep with the NCCL all-gather bit for bit, and the 1-GPU sweep to rounding, for the cyclic and the
work-balanced shard maps, at D=32 (warp per row) and D=100 (CTA per row); plus the sharded Macau feature path.
Replaces: shipping `sample_m` to every worker per half-sweep, src/sampling.jl:154-171."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, script, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, script), *map(str, args)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("D", [32, 100])
def test_sharded_sweep_on_real_peers_equals_nccl_and_one_gpu(D):
    if _ngpu() < 2:
        pytest.skip("needs at least two GPUs")
    out = _torchrun(min(_ngpu(), 4), "tools/mgpu_check.py", D)
    assert "MGPU OK" in out, out[-2000:]


def test_macau_through_the_reference_api_on_several_devices():
    if _ngpu() < 2:
        pytest.skip("needs at least two GPUs")
    out = _torchrun(2, "tools/mgpu_macau_check.py")
    assert "MGPU MACAU OK" in out, out[-2000:]


def test_macau_devices_from_a_plain_process_spawns_its_workers():
    if _ngpu() < 2:
        pytest.skip("needs at least two GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mgpu_macau_spawn.py")], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU MACAU SPAWN OK" in r.stdout, r.stdout[-2000:] + "\n" + r.stderr[-3000:]
