"""Generates the golden fixtures of tests/golden/ by RUNNING THE REFERENCE (oracle/_ref/
libsaena_ref.so = the unmodified /root/reference sources compiled by oracle/Makefile, one MPI
rank).  Run it in the build container where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

The reference ships no golden vectors of its own, and its setup is not
reproducible run to run (Lanczos start vector from std::random_device,
external/lambda_lanczos/.../lambda_lanczos.hpp:35-41), so one run is frozen here: the
hierarchy the reference's setup produced, seeded input vectors, and the outputs of the
reference's own hot-path functions on them.  Files are small (Poisson mx=12 -> 1000 rows).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref import RefOptions, RefSolver  # noqa: E402
from saena_b200.hierarchy import hierarchy_to_arrays  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def read_mtx(path):
    """MatrixMarket coordinate file -> 0-based COO as the reference's own readers take it
    (saena::matrix::read_file keeps the entries as listed: general storage, no symmetrisation)"""
    rows, cols, vals, n = [], [], [], None
    with open(path) as f:
        for line in f:
            if line.startswith("%") or not line.strip():
                continue
            t = line.split()
            if n is None:
                n = int(t[0])
                assert int(t[1]) == n
                continue
            rows.append(int(t[0]) - 1); cols.append(int(t[1]) - 1); vals.append(float(t[2]))
    return n, np.array(rows, np.int32), np.array(cols, np.int32), np.array(vals)


def band_coo(n: int, b: int):
    """pattern and values of saena::band_matrix (/root/reference/src/aux_functions2.cpp:1296-1381):
    row i, columns j in [i-b, i+b] inside [0, n), value 1/(i+j+1) (BASELINE.json configs[3] shape)"""
    i = np.repeat(np.arange(n), 2 * b + 1)
    j = i + np.tile(np.arange(-b, b + 1), n)
    keep = (j >= 0) & (j < n)
    i, j = i[keep], j[keep]
    return n, i.astype(np.int32), j.astype(np.int32), 1.0 / (i + j + 1.0)


def make(name: str, mx, opts: RefOptions, coo=None, rhs=None, with_pcg: bool = True):
    s = RefSolver.poisson(mx, opts) if coo is None else RefSolver.from_coo(*coo, rhs, opts)
    h = s.hierarchy()
    rng = np.random.default_rng(12345)
    out = {("hier." + k): v for k, v in hierarchy_to_arrays(h).items()}
    out["rhs"] = s.rhs()
    for l, lv in enumerate(h.levels):
        v = rng.uniform(-1, 1, lv.A.M)
        b = rng.uniform(-1, 1, lv.A.M)
        out[f"in.L{l}.v"] = v
        out[f"in.L{l}.b"] = b
        out[f"out.L{l}.A_matvec"] = s.matvec(l, 0, v)
        out[f"out.L{l}.residual"] = s.residual(l, v, b)
        out[f"out.L{l}.chebyshev3"] = s.smooth(l, "chebyshev", 3, v, b)
        out[f"out.L{l}.chebyshev1"] = s.smooth(l, "chebyshev", 1, v, b)
        out[f"out.L{l}.jacobi2"] = s.smooth(l, "jacobi", 2, v, b)
        out[f"out.L{l}.vcycle"] = s.vcycle(l, np.zeros(lv.A.M), b)
        out[f"out.L{l}.vcycle_jacobi_1_2"] = s.vcycle(l, v, b, pre=1, post=2, smoother="jacobi")
        if lv.P is not None:
            vc = rng.uniform(-1, 1, lv.P.n_local_cols)
            out[f"in.L{l}.vc"] = vc
            out[f"out.L{l}.P_matvec"] = s.matvec(l, 1, vc)
            out[f"out.L{l}.R_matvec"] = s.matvec(l, 2, v)
    bc = rng.uniform(-1, 1, h.coarse_n)
    out["in.coarsest.b"] = bc
    out["out.coarsest.u"] = s.coarsest_solve(bc)
    out["out.dot"] = np.array([s.dot(out["in.L0.v"], out["in.L0.b"])])
    iters = -1
    if with_pcg:
        u, iters, hist = s.solve_pcg()
        out["out.pcg.u"] = u
        out["out.pcg.iters"] = np.array([iters])
        out["out.pcg.hist"] = hist
    out["opts"] = np.array([opts.max_iter, opts.tol, opts.pre, opts.post, opts.float_level])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "levels", len(h.levels), "pcg iters", iters, "bytes", os.path.getsize(path))
    s.close()


DATA = "/root/reference/data"


def make_find_eig():
    """find_eig.npz: the reference's own Lanczos engine (LambdaLanczos via saena_object::find_eig's
    sequence, oracle/ref_harness.cpp:sref_find_eig_start) on every level of three of the matrices
    above, with seeded start vectors in place of std::random_device.  The operators of a hierarchy are
    reproducible run to run (only the eigenvalue estimates are not), so these belong to the
    hierarchies frozen in the other files."""
    rng = np.random.default_rng(777)
    out = {}
    helm = read_mtx(f"{DATA}/Helmholtz2D_CG_curved_tri/Helmholtz2D_CG_P8_Modes_curved_tri.mtx")
    for name, build in (("poisson9_cheb", lambda: RefSolver.poisson(9, RefOptions())),
                        ("poisson12_cheb", lambda: RefSolver.poisson(12, RefOptions())),
                        ("helmholtz2d_p8", lambda: RefSolver.from_coo(*helm, np.ones(helm[0]), RefOptions()))):
        s = build()
        for l, lv in enumerate(s.hierarchy().levels):
            start = rng.uniform(-1, 1, lv.A.M)
            eig, iters = s.find_eig(l, start)
            out[f"{name}.L{l}.start"] = start
            out[f"{name}.L{l}.eig"] = np.array([eig])
            out[f"{name}.L{l}.iters"] = np.array([iters])
            print(name, l, lv.A.M, eig, iters)
        s.close()
    np.savez_compressed(os.path.join(HERE, "find_eig.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["poisson12_cheb", "poisson9_cheb", "helmholtz2d_p8", "homg33", "band8_1500", "find_eig",
                             "poisson12_dense"]
    if "find_eig" in which:
        make_find_eig()
    rng = np.random.default_rng(2024)
    if "poisson12_cheb" in which:
        make("poisson12_cheb", 12, RefOptions())
    if "poisson9_cheb" in which:
        make("poisson9_cheb", 9, RefOptions())
    if "poisson12_dense" in which:
        # switch_to_dense on (data/options006_poisson.xml has it off): the coarse levels whose density exceeds
        # dense_thre = 0.1 are applied through saena_matrix_dense -- with float_level 0 that product casts the whole
        # input vector to float (src/saena_matrix_dense.cpp:262-340), which the sparse path does not
        # (dense_thre lowered to 0.05 so that level 1 of this small problem -- 500 rows, density 0.097 -- is
        # dense too: a dense level that is smoothed inside the V-cycle, not only the coarsest one)
        os.environ["SREF_SWITCH_TO_DENSE"], os.environ["SREF_DENSE_THRE"] = "1", "0.05"
        make("poisson12_dense", 12, RefOptions())
        del os.environ["SREF_SWITCH_TO_DENSE"], os.environ["SREF_DENSE_THRE"]
    if "helmholtz2d_p8" in which:
        # BASELINE.json configs[4]'s shape donor: irregular rows (24..40 non-zeros), general storage
        coo = read_mtx(f"{DATA}/Helmholtz2D_CG_curved_tri/Helmholtz2D_CG_P8_Modes_curved_tri.mtx")
        make("helmholtz2d_p8", None, RefOptions(), coo, rng.uniform(-1, 1, coo[0]))
    if "homg33" in which:
        coo = read_mtx(f"{DATA}/homg/A.mtx")
        rhs = np.loadtxt(f"{DATA}/homg/rhs.txt", skiprows=1)[:, 2]   # "i 1 value" lines under a size header
        make("homg33", None, RefOptions(), coo, rhs)
    if "band8_1500" in which:
        # BASELINE.json configs[3]'s pattern (experiments/banded.cpp), small; 1/(i+j+1) is not
        # diagonally dominant, so only the per-operator outputs are frozen, not a solve
        coo = band_coo(1500, 8)
        make("band8_1500", None, RefOptions(), coo, rng.uniform(-1, 1, coo[0]), with_pcg=False)
