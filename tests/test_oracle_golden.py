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
