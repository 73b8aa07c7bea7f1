"""ctypes bridge to oracle/_ref/libsaena_ref.so -- the UNMODIFIED reference compiled by
oracle/Makefile (TEST INFRASTRUCTURE: importable only from tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke()).

It runs the reference's own host setup (one MPI rank) and returns the finished hierarchy in
`saena_b200.hierarchy` containers, and calls the reference's own hot-path functions as the
parity oracle.  See oracle/ref_harness.cpp for the C side and the reference lines it drives.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

from saena_b200.hierarchy import (F64, I32, KIND_A, KIND_P, KIND_R, Hierarchy, Level, Operator)

_HERE = os.path.dirname(os.path.abspath(__file__))
# the ranks of `python -m oracle.mprun` (SBMPI_SIZE set) load the build against the multi-process MPI
# stand-in (make -C oracle ref_mp); everything else the one-rank build
MP_LIB_PATH = os.path.join(_HERE, "_ref", "libsaena_ref_mp.so")
LIB_PATH = (os.environ.get("SAENA_REF_LIB_PATH")     # an explicit build, e.g. the adaptor-recording one
            or (MP_LIB_PATH if os.environ.get("SBMPI_SIZE") else os.path.join(_HERE, "_ref", "libsaena_ref.so")))
REC_LIB_PATH = os.path.join(_HERE, "_ref", "libsaena_dropin_rec_mp.so")


def mp_available() -> bool:
    return os.path.exists(MP_LIB_PATH)

# field ids of sref_array (ref_harness.cpp)
F_NNZ_PER_ROW_LOCAL, F_COL_LOCAL, F_VAL_LOCAL, F_INV_DIAG, F_SPLIT, F_SPLIT_NEW = range(6)
F_ROW_REMOTE, F_VAL_REMOTE, F_NNZ_PER_COL_REMOTE, F_ENTRY_ROW, F_ENTRY_COL, F_ENTRY_VAL = range(6, 12)
F_INV_SQ_DIAG_ORIG = 12
F_VINDEX, F_SEND_PROC_RANK, F_SEND_PROC_COUNT, F_VDISPLS, F_RECV_PROC_RANK, F_RECV_PROC_COUNT, F_RDISPLS = range(13, 20)


class _Opts(ctypes.Structure):
    _fields_ = [("max_iter", ctypes.c_int), ("tol", ctypes.c_double), ("smoother", ctypes.c_char_p),
                ("pre", ctypes.c_int), ("post", ctypes.c_int), ("psmoother", ctypes.c_char_p),
                ("conn_str", ctypes.c_float), ("dynamic_levels", ctypes.c_int), ("max_level", ctypes.c_int),
                ("float_level", ctypes.c_int), ("filter_thre", ctypes.c_double), ("filter_max", ctypes.c_double),
                ("filter_start", ctypes.c_int), ("filter_rate", ctypes.c_int)]


class _LevelInfo(ctypes.Structure):
    _fields_ = [("M", ctypes.c_int), ("Mbig", ctypes.c_int), ("Nbig", ctypes.c_int),
                ("nnz_l", ctypes.c_long), ("nnz_local", ctypes.c_long), ("nnz_remote", ctypes.c_long),
                ("col_remote_size", ctypes.c_int), ("vIndexSize", ctypes.c_int), ("recvSize", ctypes.c_int),
                ("numRecvProc", ctypes.c_int), ("numSendProc", ctypes.c_int),
                ("use_double", ctypes.c_int), ("active", ctypes.c_int), ("eig_max", ctypes.c_double),
                ("use_dense", ctypes.c_int)]


@dataclass
class RefOptions:
    """The options of /root/reference/data/options006_poisson.xml (the Poisson driver's file)."""
    max_iter: int = 50
    tol: float = 1e-8
    smoother: str = "chebyshev"
    pre: int = 3
    post: int = 3
    psmoother: str = "jacobi"
    conn_str: float = 0.2
    dynamic_levels: bool = True
    max_level: int = 20
    float_level: int = 0
    filter_thre: float = 1e-12
    filter_max: float = 1e-9
    filter_start: int = 1
    filter_rate: int = 1

    def _c(self) -> _Opts:
        return _Opts(self.max_iter, self.tol, self.smoother.encode(), self.pre, self.post, self.psmoother.encode(),
                     self.conn_str, int(self.dynamic_levels), self.max_level, self.float_level, self.filter_thre,
                     self.filter_max, self.filter_start, self.filter_rate)


def write_options_xml(path: str, o: "RefOptions | None" = None) -> str:
    """an options file for the reference's drivers (attributes are read POSITIONALLY, src/saena.cpp:456-539) with
    the values of RefOptions -- data/options006_poisson.xml's, so the drivers run where /root/reference is absent"""
    o = o or RefOptions()
    attrs = [("solver_max_iter", o.max_iter), ("solver_tol", f"{o.tol:g}"), ("smoother", o.smoother),
             ("preSmooth", o.pre), ("postSmooth", o.post), ("PSmoother", o.psmoother), ("conn_str", f"{o.conn_str:g}"),
             ("dynamic_levels", int(o.dynamic_levels)), ("max_level", o.max_level), ("float_level", o.float_level),
             ("filter_thre", f"{o.filter_thre:g}"), ("filter_max", f"{o.filter_max:g}"), ("filter_start", o.filter_start),
             ("filter_rate", o.filter_rate), ("switch_to_dense", 0), ("dense_thre", "0.1"), ("dense_sz_thre", 5000),
             ("petsc", ""), ("eig", 0)]
    with open(path, "w") as f:
        f.write('<?xml version="1.0" encoding="utf-8" ?>\n<SAENA>\n    <OPTIONS\n')
        f.write("\n".join(f'\t{k}="{v}"' for k, v in attrs))
        f.write("/>\n</SAENA>\n")
    return path


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        L = ctypes.CDLL(LIB_PATH)
        L.sref_poisson_new.restype = ctypes.c_void_p
        L.sref_poisson_new_scaled.restype = ctypes.c_void_p
        L.sref_coo_new.restype = ctypes.c_void_p
        L.sref_array.restype = ctypes.c_long
        L.sref_dot.restype = ctypes.c_double
        L.sref_time_solve_pcg.restype = ctypes.c_double
        if hasattr(L, "sref_wtime"):
            L.sref_wtime.restype = ctypes.c_double
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class RefSolver:
    """A reference `saena::amg` after set_matrix + set_rhs (one MPI rank)."""

    def __init__(self, handle, opts: RefOptions):
        self._h = ctypes.c_void_p(handle)
        self.opts = opts

    @classmethod
    def poisson(cls, mx: int, opts: RefOptions | None = None, quiet: bool = True, scale: bool = False) -> "RefSolver":
        if scale:
            # the reference's own scale=true path crashes in its host code: matrix_setup() calls
            # scale_matrix(full_scale=false) (src/saena_matrix_setup.cpp:570-571), which never fills
            # inv_sq_diag_orig (:1407-1418), and set_repartition_rhs() then dereferences the empty
            # vector (src/saena_object_repart_shrink.cpp:349-350).  Every shipped driver runs
            # scale=false (experiments/Poisson.cpp:41).  The scale hooks of the solve path are
            # therefore checked CUDA-vs-oracle only (tests/test_gpu_parity.py::test_scale_hooks).
            raise NotImplementedError("the reference segfaults with scale=true (see comment)")
        opts = opts or RefOptions()
        c = opts._c()
        return cls(lib().sref_poisson_new_scaled(int(mx), ctypes.byref(c), int(quiet), 0), opts)

    @classmethod
    def from_coo(cls, n, row, col, val, rhs, opts: RefOptions | None = None, quiet: bool = True,
                 rhs_offset: int = 0) -> "RefSolver":
        """one rank: the whole matrix and rhs.  Several ranks (oracle.mprun): each rank passes its own share of the
        entries (any rows) and a contiguous block of the rhs starting at global index rhs_offset"""
        opts = opts or RefOptions()
        c = opts._c()
        row, col = np.ascontiguousarray(row, I32), np.ascontiguousarray(col, I32)
        val, rhs = np.ascontiguousarray(val, F64), np.ascontiguousarray(rhs, F64)
        if rhs_offset or len(rhs) != n:
            f = lib().sref_coo_new_part
            f.restype = ctypes.c_void_p
            return cls(f(ctypes.c_long(len(val)), _p(row), _p(col), _p(val), int(len(rhs)), int(rhs_offset), _p(rhs),
                         ctypes.byref(c), int(quiet)), opts)
        return cls(lib().sref_coo_new(int(n), ctypes.c_long(len(val)), _p(row), _p(col), _p(val), _p(rhs),
                                      ctypes.byref(c), int(quiet)), opts)

    def close(self):
        if self._h:
            lib().sref_free(self._h)
            self._h = None

    # ---- hierarchy extraction ----
    @property
    def max_level(self) -> int:
        return lib().sref_max_level(self._h)

    def _info(self, l, kind) -> _LevelInfo:
        info = _LevelInfo()
        rc = lib().sref_level_info_get(self._h, l, kind, ctypes.byref(info))
        assert rc == 0
        return info

    def _arr(self, l, kind, field, dtype):
        n = lib().sref_array(self._h, l, kind, field, None)
        assert n >= 0
        out = np.zeros(n, dtype)
        if n:
            lib().sref_array(self._h, l, kind, field, _p(out))
        return out

    # ---- several ranks (python -m oracle.mprun): ranks of a level's plans are ranks of that level's
    #      communicator, which shrinks as the levels get small; everything is translated to WORLD ranks
    @property
    def world_size(self) -> int:
        return int(lib().sref_size())

    @property
    def world_rank(self) -> int:
        return int(lib().sref_rank())

    def level_comm(self, l) -> np.ndarray:
        """world ranks of the members of level l's communicator, in its rank order ([] if not a member)"""
        out = np.zeros(self.world_size, I32)
        n = lib().sref_level_comm(self._h, int(l), _p(out))
        return out[:n].copy()

    def _operator(self, l, kind) -> Operator:
        info = self._info(l, kind)
        W, me = self.world_size, self.world_rank
        if W == 1:
            assert info.nnz_remote == 0, "one-rank reference: no remote part expected"
            return Operator(kind=kind, level=l, M=info.M, Mbig=info.Mbig, Nbig=info.Nbig, row_offset=0, col_offset=0,
                            n_local_cols=info.Nbig,
                            nnzPerRow_local=self._arr(l, kind, F_NNZ_PER_ROW_LOCAL, I32),
                            col_local=self._arr(l, kind, F_COL_LOCAL, I32),
                            val_local=self._arr(l, kind, F_VAL_LOCAL, F64),
                            use_double=bool(info.use_double), use_dense=bool(info.use_dense))
        comm = self.level_comm(l)
        if len(comm) == 0:   # not a member of this level's communicator: an empty operator
            return Operator(kind=kind, level=l, M=0, Mbig=0, Nbig=0, row_offset=0, col_offset=0, n_local_cols=0,
                            nnzPerRow_local=np.zeros(0, I32), col_local=np.zeros(0, I32), val_local=np.zeros(0, F64),
                            nnzPerProcScan=np.zeros(W + 1, np.int64), vdispls=np.zeros(W, I32), rdispls=np.zeros(W, I32),
                            use_double=bool(info.use_double), use_dense=bool(info.use_dense), nprocs=W, rank=me)
        rl = int(np.flatnonzero(comm == me)[0])
        # row / column partitions: the fine side is A's split, the coarse side splitNew (the partition R
        # writes into, before Grid::repart_u); P and R do not each fill both of their own copies
        fine = self._arr(l, KIND_A, F_SPLIT, I32)
        if kind == KIND_A:
            rsp = csp = fine
        else:
            coarse = self._arr(l, KIND_P, F_SPLIT_NEW, I32)
            if len(coarse) == 0:
                coarse = self._arr(l, KIND_R, F_SPLIT_NEW, I32)
            rsp, csp = (fine, coarse) if kind == KIND_P else (coarse, fine)

        def to_world(per_comm_rank):
            out = np.zeros(W, I32)
            if len(per_comm_rank) >= len(comm):   # (a one-rank communicator builds no plan at all)
                out[comm] = per_comm_rank[:len(comm)]
            return out

        npc = self._arr(l, kind, F_NNZ_PER_COL_REMOTE, I32)
        rpr, rpc = self._arr(l, kind, F_RECV_PROC_RANK, I32), self._arr(l, kind, F_RECV_PROC_COUNT, I32)
        rdis = self._arr(l, kind, F_RDISPLS, I32)
        # entries of the remote block per sender (world order == communicator order for contiguous blocks)
        scan = np.zeros(W + 1, np.int64)
        csum = np.concatenate(([0], np.cumsum(npc))).astype(np.int64)
        for p, c in zip(rpr, rpc):
            scan[comm[p] + 1] = csum[rdis[p] + c] - csum[rdis[p]]
        scan = np.cumsum(scan)
        return Operator(kind=kind, level=l, M=info.M, Mbig=info.Mbig, Nbig=info.Nbig,
                        row_offset=int(rsp[rl]), col_offset=int(csp[rl]), n_local_cols=int(csp[rl + 1] - csp[rl]),
                        nnzPerRow_local=self._arr(l, kind, F_NNZ_PER_ROW_LOCAL, I32),
                        col_local=self._arr(l, kind, F_COL_LOCAL, I32), val_local=self._arr(l, kind, F_VAL_LOCAL, F64),
                        row_remote=self._arr(l, kind, F_ROW_REMOTE, I32), val_remote=self._arr(l, kind, F_VAL_REMOTE, F64),
                        nnzPerCol_remote=npc, nnzPerProcScan=scan,
                        vIndex=self._arr(l, kind, F_VINDEX, I32),
                        vdispls=to_world(self._arr(l, kind, F_VDISPLS, I32)), rdispls=to_world(rdis),
                        sendProcRank=comm[self._arr(l, kind, F_SEND_PROC_RANK, I32)].astype(I32),
                        sendProcCount=self._arr(l, kind, F_SEND_PROC_COUNT, I32),
                        recvProcRank=comm[rpr].astype(I32), recvProcCount=rpc,
                        use_double=bool(info.use_double), use_dense=bool(info.use_dense), nprocs=W, rank=me)

    def _repart(self, l, which, comm):
        n = lib().sref_repart_plan(self._h, int(l), which, None, None, None)
        peer, off, cnt = np.zeros(n, I32), np.zeros(n, I32), np.zeros(n, I32)
        if n:
            lib().sref_repart_plan(self._h, int(l), which, _p(peer), _p(off), _p(cnt))
        return [(int(comm[p]), int(o), int(c)) for p, o, c in zip(peer, off, cnt)]

    def hierarchy(self) -> Hierarchy:
        """This rank's share of the hierarchy, in the layout the C ABI takes (world ranks everywhere)."""
        ml = self.max_level
        scale = bool(lib().sref_scale(self._h))
        W = self.world_size
        levels = []
        for l in range(ml + 1):
            info = self._info(l, KIND_A)
            lv = Level(level=l, A=self._operator(l, KIND_A), inv_diag=self._arr(l, KIND_A, F_INV_DIAG, F64),
                       eig_max=float(info.eig_max), active=bool(info.active))
            if l < ml:
                lv.P = self._operator(l, KIND_P)
                lv.R = self._operator(l, KIND_R)
                lv.M_coarse_old = lv.R.M
                lv.M_coarse = lv.R.M
                if W > 1:
                    mo, mn = ctypes.c_int(0), ctypes.c_int(0)
                    lib().sref_coarse_sizes(self._h, l, ctypes.byref(mo), ctypes.byref(mn))
                    lv.M_coarse_old, lv.M_coarse = mo.value, mn.value
                    comm = self.level_comm(l)
                    if len(comm):
                        lv.repart_send = self._repart(l, 0, comm)
                        lv.repart_recv = self._repart(l, 1, comm)
            if scale:
                lv.inv_sq_diag = self._arr(l, KIND_A, F_INV_SQ_DIAG_ORIG, F64)
            levels.append(lv)
        coarse_n = int(self._info(ml, KIND_A).Mbig)
        h = Hierarchy(levels=levels, scale=scale, coarse_n=coarse_n,
                      coarse_row=self._arr(ml, KIND_A, F_ENTRY_ROW, I32) if coarse_n else np.zeros(0, I32),
                      coarse_col=self._arr(ml, KIND_A, F_ENTRY_COL, I32) if coarse_n else np.zeros(0, I32),
                      coarse_val=self._arr(ml, KIND_A, F_ENTRY_VAL, F64) if coarse_n else np.zeros(0, F64))
        h.nprocs, h.rank = W, self.world_rank
        return h

    def rhs(self) -> np.ndarray:
        out = np.zeros(self._info(0, KIND_A).M, F64)
        lib().sref_rhs(self._h, _p(out))
        return out

    # ---- the reference's own hot-path functions ----
    def matvec(self, l, kind, v) -> np.ndarray:
        v = np.ascontiguousarray(v, F64)
        w = np.zeros(self._info(l, kind).M, F64)
        lib().sref_matvec(self._h, l, kind, _p(v), _p(w))
        return w

    def residual(self, l, u, rhs) -> np.ndarray:
        u, rhs = np.ascontiguousarray(u, F64), np.ascontiguousarray(rhs, F64)
        res = np.zeros_like(u)
        lib().sref_residual(self._h, l, _p(u), _p(rhs), _p(res))
        return res

    def smooth(self, l, smoother: str, iters: int, u, rhs) -> np.ndarray:
        u = np.array(u, F64, copy=True)
        rhs = np.ascontiguousarray(rhs, F64)
        lib().sref_smooth(self._h, l, int(smoother == "chebyshev"), int(iters), _p(u), _p(rhs))
        return u

    def dot(self, a, b) -> float:
        a, b = np.ascontiguousarray(a, F64), np.ascontiguousarray(b, F64)
        return float(lib().sref_dot(self._h, _p(a), _p(b), len(a)))

    def vcycle(self, l, u, rhs, pre=3, post=3, smoother="chebyshev") -> np.ndarray:
        u = np.array(u, F64, copy=True)
        rhs = np.array(rhs, F64, copy=True)
        lib().sref_vcycle(self._h, l, pre, post, int(smoother == "chebyshev"), _p(u), _p(rhs))
        return u

    def coarsest_solve(self, rhs) -> np.ndarray:
        rhs = np.array(rhs, F64, copy=True)
        u = np.zeros_like(rhs)
        lib().sref_coarsest_solve(self._h, _p(u), _p(rhs))
        return u

    def find_eig(self, l: int, start) -> tuple:
        """the reference's find_eig on level l with a caller-given Lanczos start vector -> (eig, steps);
        perturbs the level's values by the scale / scale-back round trip: use a solver of its own"""
        start = np.ascontiguousarray(start, F64)
        it = ctypes.c_int(0)
        f = lib().sref_find_eig_start
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        return float(f(self._h, int(l), _p(start), ctypes.byref(it))), it.value

    def set_direct_solver(self, name: str):
        lib().sref_set_direct_solver(self._h, int(name == "CG"))

    def coarsest_cg(self, rhs) -> np.ndarray:
        rhs = np.array(rhs, F64, copy=True)
        u = np.zeros_like(rhs)
        lib().sref_coarsest_cg(self._h, _p(u), _p(rhs))
        return u

    def solve_pcg(self, max_iter=None, tol=None, smoother=None, pre=None, post=None):
        """-> (u, iterations as reported by the reference (i+1), residual-norm history)"""
        o = self.opts
        max_iter = o.max_iter if max_iter is None else max_iter
        tol = o.tol if tol is None else tol
        smoother = o.smoother if smoother is None else smoother
        pre = o.pre if pre is None else pre
        post = o.post if post is None else post
        u = np.zeros(self._info(0, KIND_A).M, F64)
        hist = np.zeros(max_iter + 2, F64)
        n = ctypes.c_int(0)
        lib().sref_solve_pcg(self._h, int(max_iter), ctypes.c_double(tol), int(smoother == "chebyshev"), int(pre),
                             int(post), _p(u), _p(hist), len(hist), ctypes.byref(n), 1)
        hist = hist[:n.value]
        # the loop leaves with i == number of completed iterations - 1 on a break, and the
        # reference reports i+1 (saena_object_solve.cpp:2678-2682)
        return u, len(hist) - 1, hist

    def _solve_stationary(self, which, max_iter, tol, smoother, pre, post):
        u = np.zeros(self._info(0, KIND_A).M, F64)
        hist = np.zeros(max_iter + 2, F64)
        n = ctypes.c_int(0)
        f = lib().sref_solve_stationary
        f.restype = ctypes.c_int
        f(self._h, int(which), int(max_iter), ctypes.c_double(tol), int(smoother == "chebyshev"), int(pre), int(post),
          _p(u), _p(hist), len(hist), ctypes.byref(n))
        hist = hist[:n.value]
        return u, len(hist) - 1, hist

    def solve_vcycle(self, max_iter=50, tol=1e-8, smoother="chebyshev", pre=3, post=3):
        """saena_object::solve on the compiled reference -> (u, reported iterations, history)"""
        return self._solve_stationary(1, max_iter, tol, smoother, pre, post)

    def solve_smoother(self, max_iter=50, tol=1e-8, smoother="chebyshev", pre=3, post=3):
        """saena_object::solve_smoother on the compiled reference"""
        return self._solve_stationary(2, max_iter, tol, smoother, pre, post)

    def time_solve_pcg(self, reps: int) -> float:
        return float(lib().sref_time_solve_pcg(self._h, int(reps)))

    def time_matvec(self, level: int, reps: int = 5) -> float:
        """seconds per application of A_level, timed as saena_object::profile_matvecs does (collective)"""
        f = lib().sref_time_matvec
        f.restype = ctypes.c_double
        return float(f(self._h, int(level), int(reps)))
