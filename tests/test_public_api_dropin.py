"""The drop-in boundary end to end: the reference's own driver sequence (experiments/Poisson.cpp)
through the unchanged saena.hpp API, with saena::amg::solve_pCG and saena::matrix::matvec resolved
to saena_b200/adaptor/saena_b200_adaptor.cpp (GPU) and saena_object::solve_pCG run as the CPU
reference on the SAME hierarchy object in the same process.  oracle/_ref/libsaena_dropin.so is
built by `make -C oracle dropin` where /root/reference exists and travels with the snapshot."""
import ctypes
import os

import numpy as np
import pytest

from tests.util import ITER_SLACK, TOL_HIST, TOL_OP

pytestmark = [pytest.mark.gpu, pytest.mark.ref]

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libsaena_dropin.so")


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libsaena_dropin.so not built")
@pytest.mark.parametrize("mx", [18, 34])   # 34 -> 32^3: BASELINE.json configs[0]
def test_public_api_solve_pcg_on_gpu_matches_reference_cpu_solve(mx):
    L = ctypes.CDLL(LIB)
    cap = 64
    hg, hc = np.zeros(cap), np.zeros(cap)
    ig, ic, ng, nc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    du, dmv = ctypes.c_double(), ctypes.c_double()
    dp = ctypes.POINTER(ctypes.c_double)
    rc = L.dropin_poisson_check(mx, ctypes.byref(ig), ctypes.byref(ic), hg.ctypes.data_as(dp), ctypes.byref(ng),
                                hc.ctypes.data_as(dp), ctypes.byref(nc), cap, ctypes.byref(du), ctypes.byref(dmv))
    assert rc == 0
    assert abs(ig.value - ic.value) <= ITER_SLACK
    n = min(ng.value, nc.value)
    assert n >= 3
    assert np.max(np.abs(hg[:n] - hc[:n]) / hc[:n]) <= TOL_HIST
    assert hg[ng.value - 1] / hg[0] < 1e-8
    assert du.value < 1e-8
    assert dmv.value < TOL_OP


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


LIB_MP = os.path.join(os.path.dirname(LIB), "libsaena_dropin_mp.so")


@pytest.mark.skipif(not os.path.exists(LIB_MP) or _ngpu() < 2, reason="needs libsaena_dropin_mp.so and >= 2 GPUs")
@pytest.mark.parametrize("ranks,mx", [(2, 18), (4, 26)])
def test_public_api_on_several_mpi_ranks_matches_the_multirank_reference(ranks, mx):
    """the reference's driver on `ranks` MPI ranks (multi-process MPI stand-in), one GPU per rank: GPU solve through
    the public API vs the reference's own multi-rank CPU solve on the same hierarchy object"""
    import sys
    from oracle import mprun
    if _ngpu() < ranks:
        pytest.skip(f"needs {ranks} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rc, outs = mprun.run(ranks, [sys.executable, os.path.join(root, "tests", "dropin_mp_worker.py"), str(mx)], timeout=180,
                         env=dict(os.environ, PYTHONPATH=root), capture=True)
    assert rc == 0, "\n".join(o[-2000:] for o in outs)


EXE_B200 = os.path.join(os.path.dirname(LIB), "poisson_b200_mp")
EXE_CPU = os.path.join(os.path.dirname(LIB), "poisson_mp")


@pytest.mark.skipif(not (os.path.exists(EXE_B200) and os.path.exists(EXE_CPU)), reason="driver executables not built")
@pytest.mark.parametrize("ranks", [1, 2])
def test_the_reference_driver_itself_with_the_dropin_linked(ranks, tmp_path):
    """experiments/Poisson.cpp, unmodified, linked per INTEGRATION.md: its solve_pCG calls run on the GPU(s), its
    printed summary must be the CPU driver's (same iteration count, relative residual below the tolerance)"""
    import re
    import sys
    from oracle import mprun, ref
    if _ngpu() < ranks:
        pytest.skip(f"needs {ranks} GPUs")
    opts = ref.write_options_xml(str(tmp_path / "options.xml"))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rc_g, out_g = mprun.run(ranks, [EXE_B200, "34", opts], timeout=240, capture=True, env=dict(os.environ, PYTHONPATH=root))
    rc_c, out_c = mprun.run(ranks, [EXE_CPU, "34", opts], timeout=240, capture=True)
    assert rc_g == 0 and rc_c == 0, out_g[0][-2000:]
    it_g = [int(m) for m in re.findall(r"stopped at iteration\s+= (\d+)", out_g[0])]
    it_c = [int(m) for m in re.findall(r"stopped at iteration\s+= (\d+)", out_c[0])]
    rel_g = [float(m) for m in re.findall(r"relative residual\s+= ([0-9.e+-]+)", out_g[0])]
    assert it_g and it_c and abs(it_g[0] - it_c[0]) <= 1 and all(r < 1e-8 for r in rel_g)
    # the driver's own SpMV report (solver.profile_matvecs(), experiments/Poisson.cpp:262): same lines, the device's times
    mv_g = [float(m) for m in re.findall(r"matvec level 0\s*\n\s*min: ([0-9.e+-]+)", out_g[0])]
    mv_c = [float(m) for m in re.findall(r"matvec level 0\s*\n\s*min: ([0-9.e+-]+)", out_c[0])]
    assert mv_g and mv_c and 0 < mv_g[0] < mv_c[0]
