"""Freezes the reference's own stationary solvers on the Poisson mx=12 golden problem: saena_object::solve
(src/saena_object_solve.cpp:1883-2014) and saena_object::solve_smoother (:2017-2117), run through
oracle/_ref/libsaena_ref.so on a hierarchy of its own, together with that hierarchy (the setup's Lanczos start is
random, so the fixture carries the hierarchy the histories belong to).  Run where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden_stationary.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref import RefSolver  # noqa: E402
from saena_b200.hierarchy import hierarchy_to_arrays  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"vcycle": ("solve_vcycle", dict(max_iter=50, tol=1e-8, smoother="chebyshev", pre=3, post=3)),
         "smoother_cheb": ("solve_smoother", dict(max_iter=12, tol=1e-8, smoother="chebyshev", pre=3, post=3)),
         "smoother_jacobi": ("solve_smoother", dict(max_iter=7, tol=1e-8, smoother="jacobi", pre=2, post=0))}


def main():
    s = RefSolver.poisson(12)
    out = {"hier." + k: v for k, v in hierarchy_to_arrays(s.hierarchy()).items()}
    out["rhs"] = s.rhs()
    for name, (fn, kw) in CASES.items():
        u, it, hist = getattr(s, fn)(**kw)
        out[f"out.{name}.u"], out[f"out.{name}.iters"], out[f"out.{name}.hist"] = u, np.array([it]), hist
    s.close()
    np.savez_compressed(os.path.join(HERE, "poisson12_stationary.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.startswith("out.")})


if __name__ == "__main__":
    main()
