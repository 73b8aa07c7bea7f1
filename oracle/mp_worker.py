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


def check_adaptor_upload(L, s, rank, size, out):
    """The library loaded here is the reference + the drop-in adaptor + a stand-in for libsaena_b200.so that
    records what is uploaded (oracle/abi_recorder.cpp).  One solve through the public API
    (saena::amg::solve_pCG = the adaptor) makes the adaptor walk saena_object::grids on this rank; the recording
    must equal, array by array, what oracle/ref.py extracts from the same solver object."""
    import ctypes
    s.time_solve_pcg(1)
    I32, F64 = np.int32, np.float64
    info = [ctypes.c_int(0) for _ in range(5)]
    assert L.rec_info(*[ctypes.byref(x) for x in info]) == 0, "the adaptor never created a context"
    r_rank, r_n, r_levels, r_coarse_n, r_solves = (x.value for x in info)
    h = s.hierarchy()
    assert (r_rank, r_n, r_levels, r_solves) == (rank, size, len(h.levels), 1), (r_rank, r_n, r_levels, r_solves)
    L.rec_op_array.restype = L.rec_level_array.restype = L.rec_coarsest.restype = ctypes.c_long

    def arr(fn, args, dt):
        n = fn(*args, None)
        a = np.zeros(max(n, 0), dt)
        if n > 0:
            fn(*args, a.ctypes.data_as(ctypes.c_void_p))
        return a

    fields = [("nnzPerRow_local", I32), ("col_local", I32), ("val_local", F64), ("row_remote", I32), ("val_remote", F64),
              ("nnzPerCol_remote", I32), ("vIndex", I32), ("sendProcRank", I32), ("sendProcCount", I32), ("vdispls", I32),
              ("recvProcRank", I32), ("recvProcCount", I32), ("rdispls", I32)]
    checked = 0
    for l, lv in enumerate(h.levels):
        for k, op in ((0, lv.A), (1, lv.P), (2, lv.R)):
            if op is None:
                continue
            sc = (ctypes.c_long * 8)()
            L.rec_op_scalars(l, k, sc)
            want = (1, op.M, op.n_local_cols, op.col_offset, int(op.use_double), op.nnz_local, op.nnz_remote,
                    op.col_remote_size)
            assert tuple(sc) == want or (op.M == 0 and tuple(sc)[:4] == (1, 0, 0, 0)), (l, k, tuple(sc), want)
            for f, (name, dt) in enumerate(fields):
                got, ref_a = arr(L.rec_op_array, (l, k, f), dt), np.asarray(getattr(op, name), dt)
                assert got.shape == ref_a.shape and np.array_equal(got, ref_a), (l, "APR"[k], name, got[:8], ref_a[:8])
                checked += 1
        eig, mo, mn = ctypes.c_double(0), ctypes.c_int(0), ctypes.c_int(0)
        assert L.rec_level_aux(l, ctypes.byref(eig), ctypes.byref(mo), ctypes.byref(mn)) == 0
        assert np.array_equal(arr(L.rec_level_array, (l, 0), F64), lv.inv_diag), (l, "inv_diag")
        if lv.A.M:
            assert eig.value == lv.eig_max
        if lv.P is not None:
            assert (mo.value, mn.value) == (lv.M_coarse_old, lv.M_coarse), (l, mo.value, mn.value, lv.M_coarse_old, lv.M_coarse)
            if size > 1:
                assert arr(L.rec_level_array, (l, 1), I32).reshape(-1, 3).tolist() == [list(b) for b in lv.repart_send], (l, "send")
                assert arr(L.rec_level_array, (l, 2), I32).reshape(-1, 3).tolist() == [list(b) for b in lv.repart_recv], (l, "recv")
    # levels the reference applies through saena_matrix_dense (switch_to_dense): flagged, not refused
    buf = (ctypes.c_int * 64)()
    nd = L.rec_dense_levels(buf, 64)
    assert sorted(buf[:nd]) == [l for l, lv in enumerate(h.levels) if lv.A.use_dense and lv.A.M], (list(buf[:nd]),)
    # saena::amg::profile_matvecs through the public API (experiments/Poisson.cpp:262) = the adaptor's device timings:
    # one saena_b200_time_matvec of 5 applications per level, on every rank
    L.sref_profile_matvecs(s._h)
    nt = L.rec_timed_levels(buf, 64)
    assert list(buf[:nt]) == list(range(len(h.levels))), list(buf[:nt])
    assert r_coarse_n == h.coarse_n
    if h.coarse_n:
        assert np.array_equal(arr(L.rec_coarsest, (0,), I32), h.coarse_row)
        assert np.array_equal(arr(L.rec_coarsest, (1,), I32), h.coarse_col)
        assert np.array_equal(arr(L.rec_coarsest, (2,), F64), h.coarse_val)
    if os.environ.get("SAENA_MP_ADAPTOR_UPDATE"):
        # what a lazy update does (grids[0].A replaced by a matrix with other values; update1/2/3 themselves are
        # compiled out in this version of the reference): the next solve through the public API must upload again,
        # and what it uploads must be the new values
        before = arr(L.rec_op_array, (0, 0, 2), F64).copy()
        assert L.rec_inits() == 1
        L.sref_replace_A0_scaled_poisson(s._h, int(os.environ["SAENA_MP_ADAPTOR_UPDATE"]), ctypes.c_double(3.0))
        s.time_solve_pcg(1)
        assert L.rec_inits() == 2, "the adaptor kept solving on the stale device copy"
        after = arr(L.rec_op_array, (0, 0, 2), F64)
        assert after.shape == before.shape and np.allclose(after, 3.0 * before, rtol=1e-15)
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"adaptor_ok_{rank}"), "w") as f:
        f.write(f"{checked} arrays, {len(h.levels)} levels, "
                f"{sum(lv.A.M == 0 for lv in h.levels)} level(s) this rank is not a member of\n")


def main():
    what, mx, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    warmup = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    L = ref.lib()
    rank, size = L.sref_rank(), L.sref_size()
    # SAENA_MP_FLOAT_LEVEL: the options file's float_level (0 = the drivers' value: ghost values travel as float on
    # every level; >= the level count: every halo in double)
    opts = ref.RefOptions(float_level=int(os.environ.get("SAENA_MP_FLOAT_LEVEL", "0")))
    import time
    t_setup = time.perf_counter()
    if what == "poisson":
        s = ref.RefSolver.poisson(mx, opts)
    elif what == "unstructured":
        # BASELINE.json configs[4]'s synthetic shape, g = mx: every rank hands the reference its own block of
        # rows of the COO (saena::matrix::set + assemble repartition it by nnz)
        from saena_b200.sa_setup import unstructured2d_coo, unstructured2d_rhs
        n, row, col, val = unstructured2d_coo(mx)
        lo, hi = n * rank // size, n * (rank + 1) // size
        keep = (row >= lo) & (row < hi)
        s = ref.RefSolver.from_coo(n, row[keep], col[keep], val[keep], unstructured2d_rhs(n)[lo:hi], opts, rhs_offset=lo)
    else:
        raise SystemExit(f"unknown workload {what}")
    setup_s = time.perf_counter() - t_setup   # matrix generation + the reference's AMG setup on this rank
    if os.environ.get("SAENA_MP_ADAPTOR_CHECK"):
        check_adaptor_upload(L, s, rank, size, out)
        before = L.rec_destroys()
        s.close()                      # saena::amg::destroy() through the public API releases the device copy
        assert L.rec_destroys() == before + 1, "saena::amg::destroy() left the device context behind"
        L.sref_barrier()
        L.sref_finalize()
        return
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
        s.time_matvec(0, 3)
        # the reference's profile_matvecs figure for level 0 (src/saena_object.cpp:618-638 times 5 applications; 20 here,
        # after a warm-up, so that the figure is a bandwidth and not the first touches of a cache-resident vector)
        res["sec_per_matvec0"] = np.array([s.time_matvec(0, 20)])
        res["setup_s"] = np.array([setup_s])
    os.makedirs(out, exist_ok=True)
    np.savez(os.path.join(out, f"rank{rank}.npz"), **res)
    s.close()
    L.sref_barrier()
    L.sref_finalize()


if __name__ == "__main__":
    main()
