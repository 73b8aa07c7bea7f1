"""The CPU restatement (oracle/saena_oracle.c) against the golden vectors the reference wrote
(tests/golden/make_golden.py).  This is what pins the oracle on a box without /root/reference."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.util import GOLDEN, GOLDEN_ALL, Golden, check_ops_against_golden, check_pcg, check_vcycle_against_golden


class OracleImpl(Oracle):
    pass


@pytest.mark.parametrize("name", GOLDEN_ALL)
def test_oracle_ops_match_reference_golden(name):
    g = Golden(name)
    check_ops_against_golden(Oracle(g.hier), g)


@pytest.mark.parametrize("name", GOLDEN_ALL)
def test_oracle_vcycle_matches_reference_golden(name):
    g = Golden(name)
    check_vcycle_against_golden(Oracle(g.hier), g)


@pytest.mark.parametrize("name", [n for n in GOLDEN_ALL if Golden(n).has_pcg])
def test_oracle_pcg_matches_reference_golden(name):
    g = Golden(name)
    u, iters, hist = Oracle(g.hier).solve_pcg(g.rhs, g.max_iter, g.tol, "chebyshev", g.pre, g.post)
    check_pcg(iters, hist, u, int(g["out.pcg.iters"][0]), g["out.pcg.hist"], g["out.pcg.u"])
    if hist[-1] < hist[0] * g.tol:
        # the solve did what it says: ||A u - rhs|| / ||rhs|| below tol
        A = g.hier.levels[0].A.to_scipy_local()
        assert np.linalg.norm(A @ u - g.rhs) / np.linalg.norm(g.rhs) < g.tol
    else:
        assert iters == g.max_iter   # homg33: the reference runs out of iterations too


def test_oracle_stationary_vcycle_converges():
    g = Golden(GOLDEN[0])
    u, iters, hist = Oracle(g.hier).solve_vcycle(g.rhs, 50, 1e-8)
    assert hist[-1] / hist[0] < 1e-8 and iters < 50
    assert np.all(np.diff(hist) < 0)


def _find_eig_cases():
    import os
    from tests.util import GOLDEN_DIR
    d = np.load(os.path.join(GOLDEN_DIR, "find_eig.npz"))
    names = sorted({k.rsplit(".L", 1)[0] for k in d.files})
    return d, names


@pytest.mark.parametrize("name", _find_eig_cases()[1])
def test_oracle_find_eig_matches_reference_engine_golden(name):
    """the Lanczos bound of every level with the start vectors frozen next to the
    reference engine's answers (tests/golden/make_golden.py:make_find_eig); converged early on some
    levels (7..17 steps), out of steps (20) on others -- both must agree"""
    d, _ = _find_eig_cases()
    g = Golden(name)
    o = Oracle(g.hier)
    for l in range(len(g.hier.levels)):
        eig, iters = o.find_eig(l, d[f"{name}.L{l}.start"])
        assert abs(eig - float(d[f"{name}.L{l}.eig"][0])) <= 5e-9 * eig, (name, l)
        assert iters == int(d[f"{name}.L{l}.iters"][0]), (name, l)
        # and it is the bound the setup stored, up to the reference's random start vector
        assert abs(eig - g.hier.levels[l].eig_max) <= 2e-2 * eig


def test_oracle_stationary_solvers_match_the_frozen_reference_run():
    """saena_object::solve and saena_object::solve_smoother as the reference ran them on a hierarchy of its own
    (tests/golden/poisson12_stationary.npz, made by tests/golden/make_golden_stationary.py): iteration counts equal,
    histories within 1e-9 while the residual is far from cancellation (both solvers recompute r = A u - rhs every
    iteration: 1e-6 on the tail), solutions equal"""
    import os

    import numpy as np

    from saena_b200.hierarchy import hierarchy_from_arrays
    from tests.util import GOLDEN_DIR, TOL_HIST, rel
    d = np.load(os.path.join(GOLDEN_DIR, "poisson12_stationary.npz"))
    h = hierarchy_from_arrays({k[5:]: d[k] for k in d.files if k.startswith("hier.")})
    o = Oracle(h)
    cases = {"vcycle": ("solve_vcycle", dict(max_iter=50, tol=1e-8, smoother="chebyshev", pre=3, post=3)),
             "smoother_cheb": ("solve_smoother", dict(max_iter=12, tol=1e-8, smoother="chebyshev", pre=3, post=3)),
             "smoother_jacobi": ("solve_smoother", dict(max_iter=7, tol=1e-8, smoother="jacobi", pre=2, post=0))}
    for name, (fn, kw) in cases.items():
        u, it, hist = getattr(o, fn)(d["rhs"], **kw)
        want = d[f"out.{name}.hist"]
        assert it == int(d[f"out.{name}.iters"][0]) and len(hist) == len(want), name
        err = np.abs(hist - want) / want
        head = want > 1e-5 * want[0]
        assert err[head].max() <= TOL_HIST and err.max() <= 1e-6, (name, err)
        assert rel(u, d[f"out.{name}.u"]) <= 1e-9, name
