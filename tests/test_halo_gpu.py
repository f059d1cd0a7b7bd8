"""Multi-GPU parity: NCCL halo exchange (C1/C2) + all-reduce (C3) around the
action kernel, one process per GPU, against the serial oracle.  Needs >= 2
GPUs (`gpurun --gpus 2`); skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except OSError:
        return 0


@pytest.mark.parametrize("world", [2, 4])
def test_distributed_device_action(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "_halo_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("HALO_OK") == world


@pytest.mark.parametrize("world", [2])
def test_distributed_matrix_free_cg(world):
    """Config-5 style solve across GPUs: halos in every mult, NCCL all-reduce
    for the dot products, Dirichlet rows; equals the serial direct solve."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29530 + world),
           os.path.join(ROOT, "tests", "_cg_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("CG_OK") == world


@pytest.mark.parametrize("world", [2, 4])
def test_distributed_dg_advection(world):
    """Config 3 across GPUs: ghost-cell halo of q + fused owner-computes kernel."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29560 + world),
           os.path.join(ROOT, "tests", "_dg_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("DG_OK") == world
