"""The multi-rank oracle against the reference ITSELF on several MPI ranks.

The reference is compiled unmodified against a multi-process MPI stand-in (oracle/ref_shim_mp, N
processes over Unix sockets, oracle/mprun.py) and runs its real multi-rank paths: nnz-balanced row
partition, local / remote split, float halo, Grid::repart_u, coarse levels shrunk onto fewer ranks,
the coarsest level on one.  Each rank's share of the hierarchy is taken exactly as the reference laid
it out (ranks translated to world ranks) and handed to the multi-rank oracle -- the same arrays the
drop-in adaptor uploads per rank -- so this pins the multi-rank restatement, float halo included,
against the reference's own numbers: frozen in tests/golden/*_np{2,4}.npz (runs everywhere), and
live at other rank counts where oracle/_ref/libsaena_ref_mp.so exists."""
import os
import sys

import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.util import (GOLDEN_MULTIRANK, TOL_HIST, TOL_HIST_F32_HALO, MultiRankGolden, check_multirank_against_golden,
                        rel)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_apply(o):
    def apply(name, l, *a):
        if name == "A":
            return o.matvec(l, 0, a[0])
        if name == "P":
            return o.matvec(l, 1, a[0])
        if name == "R":
            return o.matvec(l, 2, a[0])
        return o.smooth(l, "chebyshev" if name == "cheb3" else "jacobi", 3 if name == "cheb3" else 2, a[0], a[1])
    return apply


def _check(g, tol_hist=TOL_HIST_F32_HALO, live_float_halo=False):
    o = Oracle(g.hiers)
    worst = check_multirank_against_golden(_oracle_apply(o), g)
    u, iters, hist = o.solve_pcg(g.rhs)
    assert iters == g.iters
    n = min(len(hist), len(g.hist))
    scale = g.hist[:n].copy()
    if live_float_halo:
        # A live float-halo run is not reproducible to better than float rounding: a ghost value on a rounding
        # boundary flips (one part in 2^24 of that value) and CG carries the absolute perturbation on while <r,r>
        # falls ~100x per iteration (measured over repeated 8-rank runs: 4e-9 ... 3e-6 of <r,r>_i, always below
        # 2e-7 of <r,r>_(i-1)).  The 1e-9 pin of the history is the double-halo run of the same test.
        scale[1:] = np.maximum(scale[1:], g.hist[:n - 1])
    assert np.max(np.abs(hist[:n] - g.hist[:n]) / scale) <= tol_hist
    assert rel(np.concatenate(u), np.concatenate(g.u)) <= 1e-8
    return worst


@pytest.mark.parametrize("name", GOLDEN_MULTIRANK)
def test_multirank_oracle_matches_the_multirank_reference_golden(name):
    g = MultiRankGolden(name)
    _check(g)
    hs = g.hiers
    # the fixture really holds the reference's multi-rank structure: remote blocks, a halo plan, and
    # (4 ranks) a coarse level that left some ranks
    assert any(h.levels[0].A.nnz_remote > 0 and len(h.levels[0].A.sendProcRank) for h in hs)
    assert not any(h.levels[0].A.use_double for h in hs)          # float_level 0: ghost values travel as float
    if name == "poisson14_np4":
        assert any(h.levels[l].A.M == 0 for h in hs for l in range(1, len(h.levels) - 1))
        assert any(lv.repart_send or lv.repart_recv for h in hs for lv in h.levels)
    assert sum(h.levels[-1].A.M > 0 for h in hs) == 1             # the coarsest level lives on one rank


@pytest.mark.ref
@pytest.mark.parametrize("float_level", [0, 100], ids=["float_halo", "double_halo"])
@pytest.mark.parametrize("ranks,mx,what", [(2, 14, "poisson"), (3, 12, "poisson"), (5, 18, "poisson"), (8, 22, "poisson"),
                                           (4, 48, "unstructured"), (6, 64, "unstructured")])
def test_multirank_oracle_matches_the_live_multirank_reference(ranks, mx, what, float_level, tmp_path):
    from oracle import mprun, ref
    if not ref.mp_available():
        pytest.skip("oracle/_ref/libsaena_ref_mp.so not built (make -C oracle ref_mp)")
    out = str(tmp_path / "mp")
    rc = mprun.run(ranks, [sys.executable, "-m", "oracle.mp_worker", what, str(mx), out], timeout=600,
                   env=dict(os.environ, SAENA_MP_DUMP="1", SAENA_MP_FLOAT_LEVEL=str(float_level), PYTHONPATH=ROOT))
    assert rc == 0
    parts = []
    for r in range(ranks):
        d = np.load(os.path.join(out, f"rank{r}.npz"))
        parts.append({k: d[k] for k in d.files})
    g = MultiRankGolden(f"live {what} np{ranks} {mx}", parts)
    if float_level:
        # every halo in double: nothing left that is not reproducible, the north_star bound applies as it stands
        assert all(op.use_double for h in g.hiers for lv in h.levels for op in (lv.A, lv.P, lv.R) if op is not None)
        _check(g, TOL_HIST)
    else:
        _check(g, TOL_HIST_F32_HALO, live_float_halo=True)


@pytest.mark.ref
def test_multirank_reference_agrees_with_the_one_rank_reference(tmp_path):
    """sanity of the stand-in itself: the reference on 4 ranks and on 1 rank solve the same system (their
    hierarchies differ -- aggregation is partition-dependent -- so the solutions agree to the solver
    tolerance, not to rounding)"""
    from oracle import mprun, ref
    if not ref.mp_available():
        pytest.skip("oracle/_ref/libsaena_ref_mp.so not built (make -C oracle ref_mp)")
    out = str(tmp_path / "mp")
    assert mprun.run(4, [sys.executable, "-m", "oracle.mp_worker", "poisson", "18", out], timeout=600,
                     env=dict(os.environ, PYTHONPATH=ROOT)) == 0
    u4 = np.concatenate([np.load(os.path.join(out, f"rank{r}.npz"))["u"] for r in range(4)])
    s = ref.RefSolver.poisson(18)
    try:
        u1, it1, _ = s.solve_pcg()
    finally:
        s.close()
    assert u4.shape == u1.shape and rel(u4, u1) < 1e-6
