"""Dense coarse operators: with switch_to_dense on, the reference's setup flags coarse levels
(density > dense_thre, Mbig <= dense_sz_thre: src/saena_object_setup2.cpp:328-329) and saena_matrix::matvec sends
them through saena_matrix_dense (include/saena_matrix.tpp:5-7, src/saena_matrix_dense.cpp:181-340).  Same matrix,
same product -- except that in float precision (float_level) the dense product casts the WHOLE input vector to
float.  The oracle restates that (oracle/saena_oracle.c:so_matvec_dense); here it is pinned against the compiled
reference, one rank and several, and against the frozen run tests/golden/poisson12_dense.npz.  The CUDA side
(saena_b200_set_operator_dense) is tested in tests/test_zz_dense_levels_gpu.py."""
import glob
import os
import sys

import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.util import GOLDEN_DENSE, TOL_HIST, TOL_OP, Golden, MultiRankGolden, rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DENSE_ENV = {"SREF_SWITCH_TO_DENSE": "1", "SREF_DENSE_THRE": "0.05"}


def test_the_fixture_holds_a_smoothed_dense_level_and_tells_the_two_products_apart():
    g = Golden(GOLDEN_DENSE[0])
    dense = [l for l, lv in enumerate(g.hier.levels) if lv.A.use_dense]
    assert dense and dense[0] < len(g.hier.levels) - 1          # a dense level inside the V-cycle, not only the coarsest
    assert not any(g.hier.levels[l].A.use_double for l in dense)   # float_level 0: the float variant of the product
    o = Oracle(g.hier)
    for l in dense:
        assert rel(o.matvec(l, 0, g[f"in.L{l}.v"]), g[f"out.L{l}.A_matvec"]) <= TOL_OP
    # applied through the sparse path the same levels miss the reference by the float cast of the local values
    for l in dense:
        g.hier.levels[l].A.use_dense = False
    o2 = Oracle(g.hier)
    for l in dense:
        assert rel(o2.matvec(l, 0, g[f"in.L{l}.v"]), g[f"out.L{l}.A_matvec"]) > 1e-9


@pytest.mark.ref
@pytest.mark.parametrize("float_level", [0, 100])
def test_oracle_matches_the_live_reference_with_dense_levels(float_level, monkeypatch):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libsaena_ref.so not built")
    for k, v in DENSE_ENV.items():
        monkeypatch.setenv(k, v)
    s = ref.RefSolver.poisson(16, ref.RefOptions(float_level=float_level))
    try:
        h = s.hierarchy()
        dense = [l for l, lv in enumerate(h.levels) if lv.A.use_dense]
        assert len(dense) >= 2
        o = Oracle(h)
        rng = np.random.default_rng(3)
        for l, lv in enumerate(h.levels):
            v, b = rng.uniform(-1, 1, lv.A.M), rng.uniform(-1, 1, lv.A.M)
            assert rel(o.matvec(l, 0, v), s.matvec(l, 0, v)) <= TOL_OP
            assert rel(o.residual(l, v, b), s.residual(l, v, b)) <= TOL_OP
            assert rel(o.smooth(l, "chebyshev", 3, v, b), s.smooth(l, "chebyshev", 3, v, b)) <= TOL_OP
            assert rel(o.smooth(l, "jacobi", 2, v, b), s.smooth(l, "jacobi", 2, v, b)) <= TOL_OP
            assert rel(o.vcycle(l, np.zeros(lv.A.M), b), s.vcycle(l, np.zeros(lv.A.M), b)) <= 1e-11
        u, it, hist = s.solve_pcg()
        uo, ito, histo = o.solve_pcg(s.rhs())
        assert ito == it
        assert np.max(np.abs(hist - histo) / hist) <= TOL_HIST
        assert rel(uo, u) <= 1e-8
    finally:
        s.close()


@pytest.mark.ref
@pytest.mark.parametrize("ranks,mx", [(2, 14), (3, 18)])
def test_multirank_oracle_matches_the_multirank_reference_with_dense_levels(ranks, mx, tmp_path):
    """several MPI ranks: the dense product goes round the ring of ranks (src/saena_matrix_dense.cpp:214-256), the
    remote blocks in double or float like the local one; halos in double so that the history pins at 1e-9"""
    from oracle import mprun, ref
    from tests.test_multirank_reference import _check
    if not ref.mp_available():
        pytest.skip("oracle/_ref/libsaena_ref_mp.so not built (make -C oracle ref_mp)")
    out = str(tmp_path / "mp")
    rc = mprun.run(ranks, [sys.executable, "-m", "oracle.mp_worker", "poisson", str(mx), out], timeout=600,
                   env=dict(os.environ, SAENA_MP_DUMP="1", SAENA_MP_FLOAT_LEVEL="100", PYTHONPATH=ROOT, **DENSE_ENV))
    assert rc == 0
    parts = []
    for r in range(ranks):
        d = np.load(os.path.join(out, f"rank{r}.npz"))
        parts.append({k: d[k] for k in d.files})
    g = MultiRankGolden(f"live dense np{ranks} {mx}", parts)
    assert any(lv.A.use_dense and lv.A.M for h in g.hiers for lv in h.levels[:-1])
    _check(g, TOL_HIST)


@pytest.mark.ref
@pytest.mark.parametrize("ranks", [1, 3])
def test_adaptor_flags_dense_levels_instead_of_refusing_them(ranks, tmp_path):
    from oracle import mprun, ref
    if not os.path.exists(ref.REC_LIB_PATH):
        pytest.skip("oracle/_ref/libsaena_dropin_rec_mp.so not built (make -C oracle dropin_rec_mp)")
    out = str(tmp_path / "rec")
    rc, outs = mprun.run(ranks, [sys.executable, "-m", "oracle.mp_worker", "poisson", "16", out], timeout=600,
                         env=dict(os.environ, SAENA_MP_ADAPTOR_CHECK="1", SAENA_REF_LIB_PATH=ref.REC_LIB_PATH,
                                  PYTHONPATH=ROOT, **DENSE_ENV), capture=True)
    assert rc == 0, "\n".join(o[-1500:] for o in outs)
    assert len(glob.glob(os.path.join(out, "adaptor_ok_*"))) == ranks
