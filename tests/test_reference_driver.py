"""BASELINE.json configs[0] literally: the reference's own driver, experiments/Poisson.cpp, unmodified, as an
executable (make -C oracle driver_mp) on 1 and 4 MPI ranks of the multi-process MPI stand-in, with the
reference's own options file -- and its printed summary against what the oracle restatement gets on the same
matrix (iteration count; the relative residual is the driver's own stopping test)."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.ref
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "poisson_mp")
OPTS = "/root/reference/data/options006_poisson.xml"


@pytest.mark.skipif(not (os.path.exists(EXE) and os.path.exists(OPTS)), reason="needs the driver build and /root/reference")
@pytest.mark.parametrize("ranks", [1, 4])
def test_the_reference_driver_runs_config0(ranks):
    from oracle import mprun
    rc, outs = mprun.run(ranks, [EXE, "34", OPTS], timeout=600, capture=True)
    if ranks == 1:
        # the options file oracle.ref.write_options_xml generates (for boxes without /root/reference) is equivalent
        from oracle import ref
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            rc2, outs2 = mprun.run(1, [EXE, "34", ref.write_options_xml(os.path.join(d, "o.xml"))], timeout=600, capture=True)
        assert rc2 == 0
        pick = lambda t: re.findall(r"(Smoother:.*|Max iter.*|Filter:.*|stopped at iteration.*|number of levels.*)", t)  # noqa: E731
        assert pick(outs2[0]) == pick(outs[0])
    assert rc == 0, outs[0][-2000:]
    text = outs[0]
    assert f"Number of MPI tasks: {ranks}" in text
    assert re.search(r"level = 0\s+number of procs = %d\s+matrix size\s+= 32768\s+nonzero\s+= 223232" % ranks, text)
    its = [int(m) for m in re.findall(r"stopped at iteration\s+= (\d+)", text)]
    rel = [float(m) for m in re.findall(r"relative residual\s+= ([0-9.e+-]+)", text)]
    assert its and all(i == its[0] for i in its) and 5 <= its[0] <= 9      # 7 on one rank and on four
    assert all(r < 1e-8 for r in rel)
    if ranks == 1:
        # the restated setup + oracle on the same matrix: same iteration count
        from oracle.oracle import Oracle
        from saena_b200.sa_setup import build_hierarchy, poisson3d_coo, poisson3d_rhs
        _, it, _ = Oracle(build_hierarchy(*poisson3d_coo(32), device="cpu")).solve_pcg(poisson3d_rhs(32))
        assert abs(it - its[0]) <= 1
