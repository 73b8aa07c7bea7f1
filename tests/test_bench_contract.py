"""bench.py's reference arm (CPU only) prints one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("ranks", [1, 3])
def test_reference_arm_prints_one_contract_line(ranks):
    # 10^3 / 16^3 unknowns: seconds instead of half a minute; ranks > 1 = the reference on several MPI ranks
    # (multi-process MPI stand-in), what the arm uses on a multi-core host
    env = dict(os.environ, SAENA_BENCH_CPU_MX="12", SAENA_BENCH_CPU_MX_MP="18", SAENA_BENCH_CPU_RANKS=str(ranks))
    if ranks > 1:
        from oracle import ref
        if not ref.mp_available():
            pytest.skip("oracle/_ref/libsaena_ref_mp.so not built (make -C oracle ref_mp)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "Munknowns/s"
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["cpu_baseline"]["cores"] == (ranks if d["cpu_baseline"]["kind"] == "reference" else 1)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] > 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, SAENA_BENCH_CPU_MX="12", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_verify_properties_through_the_oracle():
    """bench.py's SAENA_BENCH_VERIFY check (symmetry / R = P^T adjointness on every level) is implementation-agnostic:
    here through the C oracle on one rank"""
    sys.path.insert(0, ROOT)
    import bench
    from oracle.oracle import Oracle
    from tests.util import GOLDEN, Golden
    g = Golden(GOLDEN[1])
    out = bench.verify_properties(Oracle(g.hier), g.hier, 0, 1)
    assert out["ok"] and out["worst"] <= 1e-14, out
