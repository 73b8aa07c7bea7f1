"""Launches tests/multigpu_check.py under torchrun when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs on one box (gpurun --gpus 2)")
@pytest.mark.parametrize("world", [2, 4])
def test_distributed_path_matches_multirank_oracle(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29540 + world),
                          os.path.join(ROOT, "tests", "multigpu_check.py")],
                         capture_output=True, text=True, timeout=900, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    fails = [l for l in out.stdout.splitlines() if "FAILED" in l or "Error" in l or "assert" in l.lower()]
    assert out.returncode == 0, "\n".join(fails[:40]) + "\n" + out.stdout[-6000:]
    assert "MULTIGPU_OK" in out.stdout


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs on one box (gpurun --gpus 2)")
@pytest.mark.xfail(strict=False, reason="first GPU run of the reference's own multi-rank layouts (written after the "
                                        "round's GPU budget was spent); CPU side is green: test_multirank_reference.py")
@pytest.mark.parametrize("world", [2, 4])
def test_reference_multirank_layout_on_gpus(world):
    """the per-rank hierarchies exactly as the reference on `world` MPI ranks laid them out (shrunk coarse
    levels, Grid::repart_u plans, float halo: tests/golden/*_np{2,4}.npz) uploaded one rank per GPU, against
    the reference's own multi-rank outputs -- what the drop-in adaptor does in a multi-rank run"""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29560 + world),
                          os.path.join(ROOT, "tests", "multigpu_check.py")],
                         capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, MASTER_ADDR="127.0.0.1", SAENA_MG_REFERENCE_GOLDEN="1"))
    assert out.returncode == 0, out.stdout[-6000:]
    assert "the reference's own" in out.stdout
