"""CUDA path on a hierarchy with dense coarse levels (switch_to_dense): the device applies such a
level from its sparse arrays and reproduces what makes the reference's dense product different -- the float cast
of the whole input vector in float precision (saena_b200_set_operator_dense, csrc/operator.cu:sb_apply).
Against the frozen reference run tests/golden/poisson12_dense.npz and the oracle.  (File name: sorted last --
written after round 1's GPU budget was spent, its first GPU run is the round-end one.)"""
import numpy as np
import pytest

from oracle.oracle import Oracle
from saena_b200.hierarchy import KIND_A
from saena_b200.native import Context
from tests.util import (GOLDEN_DENSE, TOL_OP, Golden, check_ops_against_golden, check_pcg, check_vcycle_against_golden,
                        rel)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_DENSE)
def test_dense_levels_match_reference_golden(name):
    g = Golden(name)
    dense = [l for l, lv in enumerate(g.hier.levels) if lv.A.use_dense]
    assert dense
    ctx = Context()
    try:
        ctx.upload_hierarchy(g.hier)
        check_ops_against_golden(ctx, g)
        check_vcycle_against_golden(ctx, g)
        u, iters, hist = ctx.solve_pcg(g.rhs, g.max_iter, g.tol, "chebyshev", g.pre, g.post)
        check_pcg(iters, hist, u, int(g["out.pcg.iters"][0]), g["out.pcg.hist"], g["out.pcg.u"])
        # graph replay (second solve) and the other row mappings go through the same rounded input
        u2, it2, hist2 = ctx.solve_pcg(g.rhs, g.max_iter, g.tol, "chebyshev", g.pre, g.post)
        assert it2 == iters and np.array_equal(hist2, hist)
        for mp in (1, 8, 64, 100, 0):
            for l in dense:
                ctx.set_mapping(l, KIND_A, mp)
                assert rel(ctx.matvec(l, KIND_A, g[f"in.L{l}.v"]), g[f"out.L{l}.A_matvec"]) <= TOL_OP, (mp, l)
    finally:
        ctx.close()


def test_without_the_flag_the_same_levels_miss_the_reference():
    """control: uploaded as ordinary sparse levels the dense levels differ from the reference's dense product by the
    float cast (~1e-8) -- the flag is what carries the parity, not the tolerance"""
    g = Golden(GOLDEN_DENSE[0])
    dense = [l for l, lv in enumerate(g.hier.levels) if lv.A.use_dense]
    for l in dense:
        g.hier.levels[l].A.use_dense = False
    ctx = Context()
    try:
        ctx.upload_hierarchy(g.hier)
        o = Oracle(g.hier)
        for l in dense:
            got = ctx.matvec(l, KIND_A, g[f"in.L{l}.v"])
            assert rel(got, o.matvec(l, KIND_A, g[f"in.L{l}.v"])) <= TOL_OP      # = the sparse restatement
            assert rel(got, g[f"out.L{l}.A_matvec"]) > 1e-9                      # != the reference's dense product
    finally:
        ctx.close()


def test_mapping_autotune_keeps_parity():
    """the setup-time row-mapping autotuner (native.Context.autotune_mapping: times the neighbours of the heuristic's
    choice per operator) may change mappings, never results"""
    from tests.util import GOLDEN_EXTRA
    g = Golden(GOLDEN_EXTRA[0])     # irregular rows
    ctx = Context()
    try:
        ctx.upload_hierarchy(g.hier)
        table = ctx.autotune_mapping(reps=3)
        assert table and all(t1 <= t0 for *_, t0, t1 in table)
        for l, kind, before, after, *_ in table:
            assert ctx.get_mapping(l, kind) == after
        check_ops_against_golden(ctx, g)
        check_vcycle_against_golden(ctx, g)
    finally:
        ctx.close()


@pytest.mark.parametrize("name", ["helmholtz2d_p8", "poisson12_cheb", "band8_1500"])
def test_sorted_sliced_layout_mapping_101(name):
    """mapping 101 (sliced layout over rows sorted by length inside 256-row windows, csrc/operator.cu:build_sellp):
    every operator of the hierarchy, plain product and fused epilogues, against the reference's outputs; the
    mapping sticks (one rank: no halo) and gives the sliced mapping's result bit for bit (one lane sums a row in
    column order in both)"""
    g = Golden(name)
    ctx = Context()
    try:
        ctx.upload_hierarchy(g.hier)
        bitwise = []
        for l, lv in enumerate(g.hier.levels):
            for kind, op in ((0, lv.A), (1, lv.P), (2, lv.R)):
                if op is None:
                    continue
                x = g[f"in.L{l}.vc"] if kind == 1 else g[f"in.L{l}.v"]
                ctx.set_mapping(l, kind, 100)
                w100 = ctx.matvec(l, kind, x)
                ctx.set_mapping(l, kind, 101)
                assert ctx.get_mapping(l, kind) == 101
                w101 = ctx.matvec(l, kind, x)
                assert rel(w101, w100) <= 1e-14, (l, kind)
                bitwise.append(np.array_equal(w100, w101))
        check_ops_against_golden(ctx, g)          # all A on 101 now: SpMV, residual, Chebyshev, Jacobi; P and R
        check_vcycle_against_golden(ctx, g)
        assert all(bitwise), bitwise              # same lane, same order of the same multiply-adds
    finally:
        ctx.close()
