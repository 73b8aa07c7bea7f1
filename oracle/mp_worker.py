"""One rank of a multi-rank run of the compiled reference (TEST INFRASTRUCTURE):

    python -m oracle.mprun -n 4 python -m oracle.mp_worker poisson 14 /tmp/out [reps]

Every rank runs experiments/Poisson.cpp's sequence through oracle/ref_harness.cpp on
MPI_COMM_WORLD (the multi-process stand-in of ref_shim_mp/), solves with solve_pCG, and writes
<out>/rank<r>.npz: its rows of the solution, the residual history, the iteration count, its share of
the hierarchy (the reference's own per-rank arrays) and the wall time of `reps` timed solves."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref  # noqa: E402


def main():
    what, mx, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    warmup = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    L = ref.lib()
    rank, size = L.sref_rank(), L.sref_size()
    if what == "poisson":
        s = ref.RefSolver.poisson(mx)
    elif what == "unstructured":
        # BASELINE.json configs[4]'s synthetic shape, g = mx: every rank hands the reference its own block of
        # rows of the COO (saena::matrix::set + assemble repartition it by nnz)
        from saena_b200.sa_setup import unstructured2d_coo, unstructured2d_rhs
        n, row, col, val = unstructured2d_coo(mx)
        lo, hi = n * rank // size, n * (rank + 1) // size
        keep = (row >= lo) & (row < hi)
        s = ref.RefSolver.from_coo(n, row[keep], col[keep], val[keep], unstructured2d_rhs(n)[lo:hi], rhs_offset=lo)
    else:
        raise SystemExit(f"unknown workload {what}")
    u, iters, hist = s.solve_pcg()
    res = dict(u=u, iters=np.array([iters]), hist=hist, rank=np.array([rank]), size=np.array([size]))
    if os.environ.get("SAENA_MP_DUMP"):
        # this rank's share of the hierarchy (the reference's own arrays, ranks translated to world ranks)
        # and the reference's own operators applied to slices of seeded global vectors
        from saena_b200.hierarchy import hierarchy_to_arrays
        h = s.hierarchy()
        res.update({"hier." + k: v for k, v in hierarchy_to_arrays(h).items()})
        res["rhs"] = s.rhs()
        for l, lv in enumerate(h.levels):
            if lv.A.M == 0 and not lv.active:
                continue
            g = np.random.default_rng(1000 + l)
            v_all, b_all = g.uniform(-1, 1, lv.A.Mbig), g.uniform(-1, 1, lv.A.Mbig)
            r0 = lv.A.row_offset
            v, b = v_all[r0:r0 + lv.A.M], b_all[r0:r0 + lv.A.M]
            res[f"out.L{l}.A_matvec"] = s.matvec(l, 0, v)
            res[f"out.L{l}.chebyshev3"] = s.smooth(l, "chebyshev", 3, v, b)
            res[f"out.L{l}.jacobi2"] = s.smooth(l, "jacobi", 2, v, b)
            if lv.P is not None:
                vc_all = g.uniform(-1, 1, lv.P.Nbig)
                vc = vc_all[lv.P.col_offset:lv.P.col_offset + lv.P.n_local_cols]
                res[f"out.L{l}.P_matvec"] = s.matvec(l, 1, vc)
                res[f"out.L{l}.R_matvec"] = s.matvec(l, 2, v)
    if reps:
        for _ in range(warmup):
            s.time_solve_pcg(1)
        L.sref_barrier()   # MPI_Barrier before the timed region, as experiments/Poisson.cpp:216-246
        res["sec_per_solve"] = np.array([s.time_solve_pcg(reps) / reps])
    os.makedirs(out, exist_ok=True)
    np.savez(os.path.join(out, f"rank{rank}.npz"), **res)
    s.close()
    L.sref_barrier()
    L.sref_finalize()


if __name__ == "__main__":
    main()
