"""The multi-process MPI stand-in (oracle/ref_shim_mp/mpi_multi.c, TEST INFRASTRUCTURE) on its own: point to
point in both directions beyond the socket buffers, MPI matching order, Waitany over eagerly completed
receives, reductions (built-in, MAXLOC, user-defined), scan, gathers, alltoallv, communicator split / dup /
create / create_group, probe with MPI_ANY_SOURCE.  Needs only gcc -- this is what the multi-rank
reference build (make -C oracle ref_mp) and the multi-core CPU baseline stand on."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "ref_shim_mp")


@pytest.fixture(scope="module")
def selftest_binary(tmp_path_factory):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    exe = str(tmp_path_factory.mktemp("sbmpi") / "mpi_selftest")
    subprocess.check_call([gcc, "-O2", "-Wall", "-Wextra", "-Werror", "-I", SHIM, os.path.join(SHIM, "selftest.c"),
                           os.path.join(SHIM, "mpi_multi.c"), "-o", exe])
    return exe


@pytest.mark.parametrize("ranks", [1, 2, 5])
def test_mpi_standin_selftest(selftest_binary, ranks):
    from oracle import mprun
    rc, outs = mprun.run(ranks, [selftest_binary], timeout=120, capture=True)
    assert rc == 0, "\n".join(outs)[-3000:]
    assert f"SBMPI_SELFTEST_OK {ranks}" in outs[0]


def test_a_dying_rank_takes_the_run_down(tmp_path):
    # a rank that exits without MPI_Finalize must not leave the others waiting forever
    from oracle import mprun
    script = tmp_path / "die.py"
    script.write_text("import os, sys, time\n"
                      "if os.environ['SBMPI_RANK'] == '1': sys.exit(7)\n"
                      "time.sleep(60)\n")
    rc = mprun.run(3, [sys.executable, str(script)], timeout=30)
    assert rc == 7
