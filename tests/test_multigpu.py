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


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs on one box (gpurun --gpus 2)")
@pytest.mark.parametrize("world,args", [(2, ["poisson", "40", "double"]), (2, ["poisson", "48"]), (4, ["unstructured", "300", "double"])])
def test_distributed_setup_on_gpus_feeds_the_library(world, args):
    """saena_b200/sa_setup_dist.py over NCCL (what bench.py --n 512 uses on 8 GPUs) = the one-process setup, and the
    CUDA library's solve on the shares it produces = the multi-rank oracle's (tests/dist_setup_check.py, DSC_GPU=1).
    Written after round 1's GPU budget was spent: CPU side (gloo) is green in tests/test_dist_setup.py."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29560 + world),
                          os.path.join(ROOT, "tests", "dist_setup_check.py"), *args],
                         capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, MASTER_ADDR="127.0.0.1", DSC_GPU="1", DSC_AGG_BELOW="400"))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-6000:]
    assert "DIST_SETUP_GPU_OK" in out.stdout and "DIST_SETUP_OK" in out.stdout
