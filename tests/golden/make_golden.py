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


def make(name: str, mx: int, opts: RefOptions):
    s = RefSolver.poisson(mx, opts)
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
    u, iters, hist = s.solve_pcg()
    out["out.pcg.u"] = u
    out["out.pcg.iters"] = np.array([iters])
    out["out.pcg.hist"] = hist
    out["opts"] = np.array([opts.max_iter, opts.tol, opts.pre, opts.post, opts.float_level])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "levels", len(h.levels), "pcg iters", iters, "bytes", os.path.getsize(path))
    s.close()


if __name__ == "__main__":
    make("poisson12_cheb", 12, RefOptions())
    make("poisson9_cheb", 9, RefOptions())
