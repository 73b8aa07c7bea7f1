"""CUDA path vs the C oracle at sizes between the goldens (<= 32^3) and BASELINE.json's 256^3: the Poisson
hierarchy of the bench generator (saena_b200/sa_setup.py, built on the GPU in seconds) at 96^3 and 128^3 --
0.9 / 2.1 M rows on level 0, hundreds of thousands of rows on the sliced levels 1-2, 1 000+-entry rows on the
row-group levels.  Every operator, fused smoother sweep and transfer within 1e-12, V-cycle within 1e-11, the PCG
iteration count +-1 and the WHOLE residual history within 1e-9 of the oracle's (north_star tolerances).
The oracle needs ~10 s (96^3) / ~30 s (128^3) of one host core per solve."""
import os

import numpy as np
import pytest

from tests.util import TOL_HIST, TOL_OP, check_pcg, rel

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(1500)
@pytest.mark.parametrize("n", [96, 128])
def test_midsize_poisson_matches_oracle(n):
    import torch

    from oracle.oracle import Oracle
    from saena_b200 import sa_setup
    from saena_b200.hierarchy import KIND_A, KIND_P, KIND_R
    from saena_b200.native import Context

    n = int(os.environ.get(f"SAENA_TEST_MID_N{n}", n))
    dh = sa_setup.build_device_hierarchy(*sa_setup.poisson3d_coo(n), device="cuda")
    hier = dh.to_rank(0, 1)
    del dh
    torch.cuda.empty_cache()
    rhs = sa_setup.poisson3d_rhs(n)
    o = Oracle(hier)
    ctx = Context()
    rng = np.random.default_rng(96)
    worst = {}
    try:
        ctx.upload_hierarchy(hier)
        for l, lv in enumerate(hier.levels):
            M = lv.A.M
            v, b = rng.uniform(-1, 1, M), rng.uniform(-1, 1, M)
            worst[f"L{l}.A"] = rel(ctx.matvec(l, KIND_A, v), o.matvec(l, KIND_A, v))
            worst[f"L{l}.residual"] = rel(ctx.residual(l, v, b), o.residual(l, v, b))
            worst[f"L{l}.cheb3"] = rel(ctx.smooth(l, "chebyshev", 3, v, b), o.smooth(l, "chebyshev", 3, v, b))
            worst[f"L{l}.jacobi2"] = rel(ctx.smooth(l, "jacobi", 2, v, b), o.smooth(l, "jacobi", 2, v, b))
            if lv.P is not None:
                e = rng.uniform(-1, 1, lv.P.n_local_cols)
                worst[f"L{l}.P"] = rel(ctx.matvec(l, KIND_P, e), o.matvec(l, KIND_P, e))
                worst[f"L{l}.R"] = rel(ctx.matvec(l, KIND_R, v), o.matvec(l, KIND_R, v))
        bad = {k: e for k, e in worst.items() if not e <= TOL_OP}
        assert not bad, f"n={n}: beyond {TOL_OP:g}: {bad}"
        M0 = hier.levels[0].A.M
        b = rng.uniform(-1, 1, M0)
        ev = rel(ctx.vcycle(0, np.zeros(M0), b), o.vcycle(0, np.zeros(M0), b))
        assert ev <= 1e-11, ("vcycle", ev)
        u, iters, hist = ctx.solve_pcg(rhs, 50, 1e-8, "chebyshev", 3, 3)
        u_o, it_o, h_o = o.solve_pcg(rhs, 50, 1e-8, "chebyshev", 3, 3)
        err = check_pcg(iters, np.asarray(hist), u, it_o, np.asarray(h_o), u_o, tol_hist=TOL_HIST)
        print(f"n={n}: {len(hier.levels)} levels, worst op {max(worst.values()):.2e}, vcycle {ev:.2e}, "
              f"pcg {iters} iterations (oracle {it_o}), history err {err:.2e}, "
              f"hist/hist0 {[float(f'{x:.3e}') for x in np.asarray(hist) / hist[0]]}")
        assert ctx.launch_count() > 0
    finally:
        ctx.close()
