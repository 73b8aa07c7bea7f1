"""Shared helpers of the parity tests.

`impl` below is anything with the methods matvec / residual / smooth / vcycle / coarsest_solve /
dot and a PCG entry point: the C oracle (oracle.oracle.Oracle), the compiled reference
(oracle.ref.RefSolver) or the CUDA path (saena_b200.native.Context) -- the three expose the same
names on purpose, so one checker serves "oracle vs golden", "oracle vs reference" and
"CUDA vs oracle/golden".
"""
import os

import numpy as np

from saena_b200.hierarchy import hierarchy_from_arrays

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = ["poisson9_cheb", "poisson12_cheb"]
# other matrix shapes frozen from the reference (tests/golden/make_golden.py): the irregular-row
# Helmholtz matrix BASELINE.json configs[4] takes its shape from (data/Helmholtz2D_CG_curved_tri),
# data/homg/A.mtx with its own rhs (PCG does not converge on it in the reference either: the whole
# 50-iteration history is the fixture), and the band pattern of configs[3] (experiments/banded.cpp;
# per-operator outputs only, 1/(i+j+1) is not a system one solves)
GOLDEN_EXTRA = ["helmholtz2d_p8", "homg33", "band8_1500"]
# switch_to_dense on: coarse levels applied through saena_matrix_dense (Operator.use_dense), float precision
GOLDEN_DENSE = ["poisson12_dense"]
GOLDEN_ALL = GOLDEN + GOLDEN_EXTRA + GOLDEN_DENSE

# tolerances of BASELINE.json's north_star
TOL_OP = 1e-12       # each SpMV, smoother sweep and transfer: relative error in the 2-norm
TOL_HIST = 1e-9      # residual-norm history, relative, per iteration
ITER_SLACK = 1       # iteration count +-1


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / nb) if nb > 0 else float(np.linalg.norm(a))


class Golden:
    def __init__(self, name):
        self.name = name
        self.d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.hier = hierarchy_from_arrays({k[5:]: self.d[k] for k in self.d.files if k.startswith("hier.")})
        self.rhs = self.d["rhs"]
        self.max_iter, self.tol, self.pre, self.post = (int(self.d["opts"][0]), float(self.d["opts"][1]),
                                                        int(self.d["opts"][2]), int(self.d["opts"][3]))

    def __getitem__(self, k):
        return self.d[k]

    @property
    def has_pcg(self):
        return "out.pcg.hist" in self.d.files


def check_ops_against_golden(impl, g: Golden, tol=TOL_OP):
    """every per-operator function on the golden inputs vs what the reference returned"""
    h = g.hier
    worst = {}
    for l, lv in enumerate(h.levels):
        v, b = g[f"in.L{l}.v"], g[f"in.L{l}.b"]
        worst[f"L{l}.A"] = rel(impl.matvec(l, 0, v), g[f"out.L{l}.A_matvec"])
        worst[f"L{l}.residual"] = rel(impl.residual(l, v, b), g[f"out.L{l}.residual"])
        worst[f"L{l}.cheb3"] = rel(impl.smooth(l, "chebyshev", 3, v, b), g[f"out.L{l}.chebyshev3"])
        worst[f"L{l}.cheb1"] = rel(impl.smooth(l, "chebyshev", 1, v, b), g[f"out.L{l}.chebyshev1"])
        worst[f"L{l}.jacobi2"] = rel(impl.smooth(l, "jacobi", 2, v, b), g[f"out.L{l}.jacobi2"])
        if lv.P is not None:
            worst[f"L{l}.P"] = rel(impl.matvec(l, 1, g[f"in.L{l}.vc"]), g[f"out.L{l}.P_matvec"])
            worst[f"L{l}.R"] = rel(impl.matvec(l, 2, v), g[f"out.L{l}.R_matvec"])
    worst["coarsest"] = rel(impl.coarsest_solve(g["in.coarsest.b"]), g["out.coarsest.u"])
    d = impl.dot(g["in.L0.v"], g["in.L0.b"])
    worst["dot"] = abs(d - float(g["out.dot"][0])) / abs(float(g["out.dot"][0]))
    bad = {k: e for k, e in worst.items() if not e <= tol}
    assert not bad, f"{g.name}: beyond {tol:g}: {bad}"
    return worst


def check_vcycle_against_golden(impl, g: Golden, tol=1e-11):
    # a V-cycle chains ~10 operator applications per level; allow the per-op bound to accumulate
    h = g.hier
    worst = {}
    for l, lv in enumerate(h.levels):
        v, b = g[f"in.L{l}.v"], g[f"in.L{l}.b"]
        worst[f"L{l}.vcycle"] = rel(impl.vcycle(l, np.zeros(lv.A.M), b), g[f"out.L{l}.vcycle"])
        worst[f"L{l}.vcycle_jac"] = rel(impl.vcycle(l, v, b, pre=1, post=2, smoother="jacobi"),
                                         g[f"out.L{l}.vcycle_jacobi_1_2"])
    bad = {k: e for k, e in worst.items() if not e <= tol}
    assert not bad, f"{g.name}: beyond {tol:g}: {bad}"
    return worst


def check_pcg(iters, hist, u, ref_iters, ref_hist, ref_u, tol_hist=TOL_HIST):
    assert abs(iters - ref_iters) <= ITER_SLACK, (iters, ref_iters)
    n = min(len(hist), len(ref_hist))
    err = np.abs(hist[:n] - ref_hist[:n]) / ref_hist[:n]
    assert err.max() <= tol_hist, f"residual history differs: {err}"
    if iters == ref_iters:
        assert rel(u, ref_u) <= 1e-8
    return float(err.max())


# ---- multi-rank goldens: the reference itself on N MPI ranks (tests/golden/make_golden_multirank.py) ----
GOLDEN_MULTIRANK = ["poisson10_np2", "poisson14_np4", "unstructured40_np3"]
TOL_HIST_F32_HALO = 1e-6   # float_level 0: see DESIGN.md section 2 (float rounding flips of single ghost values)


class MultiRankGolden:
    """per-rank hierarchies exactly as the reference laid them out + its own outputs"""

    def __init__(self, name=None, parts=None):
        if parts is None:
            d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
            n = int(d["ranks"][0])
            parts = [{k[len(f"r{r}."):]: d[k] for k in d.files if k.startswith(f"r{r}.")} for r in range(n)]
        self.name, self.parts, self.nranks = name, parts, len(parts)
        self.hiers = [hierarchy_from_arrays({k[5:]: p[k] for k in p if k.startswith("hier.")}) for p in parts]
        self.rhs = [p["rhs"] for p in parts]
        self.iters = int(parts[0]["iters"][0])
        self.hist = parts[0]["hist"]
        self.u = [p["u"] for p in parts]

    def inputs(self, l):
        """the seeded global vectors of level l, sliced per rank (what oracle/mp_worker.py fed the reference)"""
        hs = self.hiers
        g = np.random.default_rng(1000 + l)
        Mbig = max(h.levels[l].A.Mbig for h in hs)
        v_all, b_all = g.uniform(-1, 1, Mbig), g.uniform(-1, 1, Mbig)
        sl = [(h.levels[l].A.row_offset, h.levels[l].A.M) for h in hs]
        v, b = [v_all[o:o + m] for o, m in sl], [b_all[o:o + m] for o, m in sl]
        vc = None
        if hs[0].levels[l].P is not None:
            vc_all = g.uniform(-1, 1, max(h.levels[l].P.Nbig for h in hs))
            vc = [vc_all[h.levels[l].P.col_offset:h.levels[l].P.col_offset + h.levels[l].P.n_local_cols] for h in hs]
        return v, b, vc

    def want(self, l, key):
        return [p.get(f"out.L{l}.{key}", np.zeros(0)) for p in self.parts]


def check_multirank_against_golden(apply, g: MultiRankGolden, tol=TOL_OP):
    """apply(op_name, level, *per-rank inputs) -> per-rank outputs; op_name in A / cheb3 / jac2 / P / R"""
    worst = {}
    cat = np.concatenate
    for l in range(len(g.hiers[0].levels)):
        v, b, vc = g.inputs(l)
        worst[f"L{l}.A"] = rel(cat(apply("A", l, v)), cat(g.want(l, "A_matvec")))
        worst[f"L{l}.cheb3"] = rel(cat(apply("cheb3", l, v, b)), cat(g.want(l, "chebyshev3")))
        worst[f"L{l}.jac2"] = rel(cat(apply("jac2", l, v, b)), cat(g.want(l, "jacobi2")))
        if vc is not None:
            worst[f"L{l}.P"] = rel(cat(apply("P", l, vc)), cat(g.want(l, "P_matvec")))
            worst[f"L{l}.R"] = rel(cat(apply("R", l, v)), cat(g.want(l, "R_matvec")))
    bad = {k: e for k, e in worst.items() if not e <= tol}
    assert not bad, f"{g.name}: beyond {tol:g}: {bad}"
    return worst
