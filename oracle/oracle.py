"""ctypes bridge to oracle/libsaena_oracle.so -- the CPU restatement of the reference's solve
path (TEST INFRASTRUCTURE: importable only from tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke(); see saena_oracle.h).

All ranks of a partitioned hierarchy are emulated in one process: pass the list of per-rank
`Hierarchy` objects (`saena_b200.hierarchy.partition_hierarchy`) or a single one.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Sequence

import numpy as np

from saena_b200.hierarchy import F64, I32, I64, Hierarchy, Operator

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsaena_oracle.so")

c_int_p = ctypes.POINTER(ctypes.c_int)
c_long_p = ctypes.POINTER(ctypes.c_long)
c_double_p = ctypes.POINTER(ctypes.c_double)


class _Op(ctypes.Structure):
    _fields_ = [("M", ctypes.c_int), ("n_local_cols", ctypes.c_int), ("col_offset", ctypes.c_int),
                ("nnz_local", ctypes.c_long), ("nnzPerRow_local", c_int_p), ("col_local", c_int_p),
                ("val_local", c_double_p), ("nnz_remote", ctypes.c_long), ("col_remote_size", ctypes.c_int),
                ("row_remote", c_int_p), ("val_remote", c_double_p), ("nnzPerCol_remote", c_int_p),
                ("nnzPerProcScan", c_long_p), ("vIndexSize", ctypes.c_int), ("vIndex", c_int_p),
                ("vdispls", c_int_p), ("rdispls", c_int_p), ("numSendProc", ctypes.c_int),
                ("sendProcRank", c_int_p), ("sendProcCount", c_int_p), ("numRecvProc", ctypes.c_int),
                ("recvProcRank", c_int_p), ("recvProcCount", c_int_p), ("use_double", ctypes.c_int),
                ("use_dense", ctypes.c_int)]


class _Block(ctypes.Structure):
    _fields_ = [("peer", ctypes.c_int), ("offset", ctypes.c_int), ("count", ctypes.c_int)]


class _Level(ctypes.Structure):
    _fields_ = [("A", _Op), ("P", _Op), ("R", _Op), ("inv_diag", c_double_p), ("inv_sq_diag", c_double_p),
                ("eig_max", ctypes.c_double),
                ("M_coarse_old", ctypes.c_int), ("M_coarse", ctypes.c_int), ("n_repart_send", ctypes.c_int),
                ("n_repart_recv", ctypes.c_int), ("repart_send", ctypes.POINTER(_Block)),
                ("repart_recv", ctypes.POINTER(_Block))]


class _Hier(ctypes.Structure):
    _fields_ = [("nranks", ctypes.c_int), ("nlevels", ctypes.c_int), ("level", ctypes.POINTER(_Level)),
                ("coarse_n", ctypes.c_int), ("coarse_dense", c_double_p), ("scale", ctypes.c_int),
                ("coarsest_cg", ctypes.c_int)]


def build() -> str:
    """Compile the restatement (gcc, a second or two).  Building the checker is not using it."""
    src = os.path.join(_HERE, "saena_oracle.c")
    if (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "saena_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.so_dot.restype = ctypes.c_double
        _lib = L
    return _lib


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _lp(a):
    return a.ctypes.data_as(c_long_p)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _vecs(vs: Sequence[np.ndarray]):
    """list of float64 arrays -> (double*[]) ; keeps the arrays alive through the return value"""
    arrs = [np.ascontiguousarray(v, F64) for v in vs]
    ptrs = (c_double_p * len(arrs))(*[_dp(a) for a in arrs])
    return arrs, ptrs


def _largest_tridiag_eig(alpha, beta, rel_eps):
    """largest eigenvalue of the symmetric tridiagonal matrix (diagonal alpha, off-diagonal beta) by
    bisection on the Sturm count -- find_mth_eigenvalue / num_of_eigs_smaller_than,
    /root/reference/external/lambda_lanczos/include/lambda_lanczos/lambda_lanczos.hpp:303-377"""
    m = len(alpha)

    def count_below(c):
        cnt, q = 0, 1.0
        for i in range(m):
            q = alpha[i] - c - (beta[i - 1] ** 2 / q if i else 0.0)
            if q < 0:
                cnt += 1
            if q == 0:
                q = 1e-15
        return cnt

    r = max(abs(alpha[i]) + (abs(beta[i - 1]) if i else 0.0) + (abs(beta[i]) if i + 1 < m else 0.0) for i in range(m))
    lo, hi, pmid = -r, r, None
    while hi - lo > min(abs(lo), abs(hi)) * rel_eps:
        mid = 0.5 * (lo + hi)
        if count_below(mid) >= m:
            hi = mid
        else:
            lo = mid
        if mid == pmid:
            break
        pmid = mid
    return 0.5 * (lo + hi)


class Oracle:
    """The hierarchy of every rank, laid out for the C oracle."""

    def __init__(self, hier: Hierarchy | List[Hierarchy], coarsest_cg: bool = False):
        self.ranks: List[Hierarchy] = hier if isinstance(hier, list) else [hier]
        self.nranks = len(self.ranks)
        self.nlevels = len(self.ranks[0].levels)
        self._keep = []  # numpy arrays referenced from the C structs
        self._levels = (_Level * (self.nranks * self.nlevels))()
        for l in range(self.nlevels):
            for r in range(self.nranks):
                lv = self.ranks[r].levels[l]
                c = self._levels[l * self.nranks + r]
                c.A = self._op(lv.A)
                if lv.P is not None:
                    c.P = self._op(lv.P)
                    c.R = self._op(lv.R)
                inv = np.ascontiguousarray(lv.inv_diag, F64)
                self._keep.append(inv)
                c.inv_diag = _dp(inv)
                if lv.inv_sq_diag is not None:
                    isq = np.ascontiguousarray(lv.inv_sq_diag, F64)
                    self._keep.append(isq)
                    c.inv_sq_diag = _dp(isq)
                c.eig_max = lv.eig_max
                c.M_coarse_old, c.M_coarse = lv.M_coarse_old, lv.M_coarse
                for name, blocks in (("repart_send", lv.repart_send), ("repart_recv", lv.repart_recv)):
                    arr = (_Block * max(len(blocks), 1))(*[_Block(*b) for b in blocks])
                    self._keep.append(arr)
                    setattr(c, name, arr)
                    setattr(c, "n_" + name, len(blocks))
        h0 = self.ranks[0]
        dense = np.zeros((h0.coarse_n, h0.coarse_n), F64)
        np.add.at(dense, (h0.coarse_row, h0.coarse_col), h0.coarse_val)
        self.coarse_dense = dense
        self._h = _Hier(self.nranks, self.nlevels, self._levels, h0.coarse_n, _dp(dense), int(h0.scale),
                        int(coarsest_cg))

    def _op(self, op: Operator) -> _Op:
        a = dict(nnzPerRow_local=np.ascontiguousarray(op.nnzPerRow_local, I32),
                 col_local=np.ascontiguousarray(op.col_local, I32), val_local=np.ascontiguousarray(op.val_local, F64),
                 row_remote=np.ascontiguousarray(op.row_remote, I32), val_remote=np.ascontiguousarray(op.val_remote, F64),
                 nnzPerCol_remote=np.ascontiguousarray(op.nnzPerCol_remote, I32),
                 nnzPerProcScan=np.ascontiguousarray(op.nnzPerProcScan, I64), vIndex=np.ascontiguousarray(op.vIndex, I32),
                 vdispls=np.ascontiguousarray(op.vdispls, I32), rdispls=np.ascontiguousarray(op.rdispls, I32),
                 sendProcRank=np.ascontiguousarray(op.sendProcRank, I32),
                 sendProcCount=np.ascontiguousarray(op.sendProcCount, I32),
                 recvProcRank=np.ascontiguousarray(op.recvProcRank, I32),
                 recvProcCount=np.ascontiguousarray(op.recvProcCount, I32))
        self._keep.append(a)
        return _Op(op.M, op.n_local_cols, op.col_offset, op.nnz_local, _ip(a["nnzPerRow_local"]), _ip(a["col_local"]),
                   _dp(a["val_local"]), op.nnz_remote, op.col_remote_size, _ip(a["row_remote"]), _dp(a["val_remote"]),
                   _ip(a["nnzPerCol_remote"]), _lp(a["nnzPerProcScan"]), op.vIndexSize, _ip(a["vIndex"]),
                   _ip(a["vdispls"]), _ip(a["rdispls"]), len(op.sendProcRank), _ip(a["sendProcRank"]),
                   _ip(a["sendProcCount"]), len(op.recvProcRank), _ip(a["recvProcRank"]), _ip(a["recvProcCount"]),
                   int(op.use_double), int(op.use_dense))

    # ---- helpers ----
    def _ops(self, l, kind):
        name = {0: "A", 1: "P", 2: "R"}[kind]
        ptrs = (ctypes.POINTER(_Op) * self.nranks)()
        for r in range(self.nranks):
            ptrs[r] = ctypes.pointer(getattr(self._levels[l * self.nranks + r], name))
        return ptrs

    def _lvs(self, l):
        ptrs = (ctypes.POINTER(_Level) * self.nranks)()
        for r in range(self.nranks):
            ptrs[r] = ctypes.pointer(self._levels[l * self.nranks + r])
        return ptrs

    def _rows(self, l, kind=0):
        name = {0: "A", 1: "P", 2: "R"}[kind]
        return [getattr(self.ranks[r].levels[l], name).M for r in range(self.nranks)]

    def _wrap(self, v):
        """accept one array (1 rank) or a list of per-rank arrays"""
        return [v] if isinstance(v, np.ndarray) else list(v)

    def _unwrap(self, vs, like):
        return vs[0] if isinstance(like, np.ndarray) else vs

    # ---- the restated functions ----
    def matvec(self, l, kind, v):
        vin, vp = _vecs(self._wrap(v))
        wout, wp = _vecs([np.zeros(m) for m in self._rows(l, kind)])
        lib().so_matvec(self._ops(l, kind), self.nranks, vp, wp)
        return self._unwrap(wout, v)

    def residual(self, l, u, rhs):
        uin, up = _vecs(self._wrap(u))
        rin, rp = _vecs(self._wrap(rhs))
        out, op = _vecs([np.zeros(m) for m in self._rows(l)])
        lib().so_residual(self._ops(l, 0), self.nranks, up, rp, op)
        return self._unwrap(out, u)

    def smooth(self, l, smoother, iters, u, rhs):
        uin, up = _vecs([np.array(x, F64, copy=True) for x in self._wrap(u)])
        rin, rp = _vecs(self._wrap(rhs))
        fn = lib().so_chebyshev if smoother == "chebyshev" else lib().so_jacobi
        fn(self._lvs(l), self.nranks, int(iters), up, rp)
        return self._unwrap(uin, u)

    def find_eig(self, l, start, max_iter=20, eps=1e-8):
        """saena_object::find_eig (/root/reference/src/saena_object.cpp:572-590) restated: Lanczos on
        D^-1/2 A D^-1/2 as LambdaLanczos::run does it (external/lambda_lanczos/.../lambda_lanczos.hpp:170-260:
        u_0 = 0 dummy, full modified Gram-Schmidt against every earlier vector, largest Ritz value by
        bisection on the Sturm count :303-377, stop at a relative change below eps), start vector given
        by the caller (one array per rank, or one array) -> (1.0001 * eigenvalue, steps).  Pinned against
        the reference's own engine with the same start vector: tests/test_oracle_vs_reference.py."""
        parts = self._wrap(start)
        sizes = [len(p) for p in parts]
        off = np.concatenate(([0], np.cumsum(sizes)))
        split = lambda v: [v[off[r]:off[r + 1]] for r in range(self.nranks)]  # noqa: E731
        s = np.sqrt(np.abs(np.concatenate([self.ranks[r].levels[l].inv_diag for r in range(self.nranks)])))
        mv = lambda v: s * np.concatenate(self._wrap(self.matvec(l, 0, split(s * v) if self.nranks > 1 else s * v)))  # noqa: E731
        uk = np.concatenate(parts).astype(F64)
        uk = uk / np.sqrt(uk @ uk)
        u = [np.zeros_like(uk), uk]
        alpha, beta = [], []
        betak, ev, pev, itern = 0.0, 0.0, np.finfo(F64).max, max_iter
        for k in range(1, max_iter + 1):
            vk = mv(u[k])
            alphak = float(u[k] @ vk)
            alpha.append(alphak)
            nxt = vk - betak * u[k - 1] - alphak * u[k]
            for uj in u:
                nxt = nxt - float(uj @ nxt) * uj
            betak = float(np.sqrt(nxt @ nxt))
            beta.append(betak)
            ev = _largest_tridiag_eig(alpha, beta[:-1], eps * 0.1)
            if betak < 1e-16:
                itern = k
                break
            u.append(nxt / betak)
            if abs(ev - pev) < min(abs(ev), abs(pev)) * eps:
                itern = k
                break
            pev = ev
        return 1.0001 * ev, itern

    def dot(self, a, b):
        ain, ap = _vecs(self._wrap(a))
        bin_, bp = _vecs(self._wrap(b))
        M = np.array([len(x) for x in ain], I32)
        return float(lib().so_dot(self.nranks, _ip(M), ap, bp))

    def coarsest_solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, F64)
        u = np.zeros_like(rhs)
        lib().so_coarsest_solve(ctypes.byref(self._h), _dp(rhs), _dp(u))
        return u

    def coarsest_cg(self, rhs, u0=None):
        rhs = np.ascontiguousarray(rhs, F64)
        u = np.zeros_like(rhs) if u0 is None else np.array(u0, F64, copy=True)
        lib().so_coarsest_cg(ctypes.byref(self._h), _dp(rhs), _dp(u))
        return u

    def vcycle(self, l, u, rhs, pre=3, post=3, smoother="chebyshev"):
        uin, up = _vecs([np.array(x, F64, copy=True) for x in self._wrap(u)])
        rin, rp = _vecs([np.array(x, F64, copy=True) for x in self._wrap(rhs)])
        lib().so_vcycle(ctypes.byref(self._h), int(l), int(smoother == "chebyshev"), int(pre), int(post), up, rp)
        return self._unwrap(uin, u)

    def _solve(self, fn, rhs, max_iter, tol, smoother, pre, post):
        rin, rp = _vecs(self._wrap(rhs))
        uout, up = _vecs([np.zeros(m) for m in self._rows(0)])
        hist = np.zeros(max_iter + 2, F64)
        n = ctypes.c_int(0)
        iters = fn(ctypes.byref(self._h), rp, up, int(max_iter), ctypes.c_double(tol),
                   int(smoother == "chebyshev"), int(pre), int(post), _dp(hist), len(hist), ctypes.byref(n))
        return self._unwrap(uout, rhs), int(iters), hist[:n.value]

    def solve_pcg(self, rhs, max_iter=50, tol=1e-8, smoother="chebyshev", pre=3, post=3):
        return self._solve(lib().so_solve_pcg, rhs, max_iter, tol, smoother, pre, post)

    def solve_vcycle(self, rhs, max_iter=50, tol=1e-8, smoother="chebyshev", pre=3, post=3):
        return self._solve(lib().so_solve_vcycle, rhs, max_iter, tol, smoother, pre, post)

    def solve_smoother(self, rhs, max_iter=50, tol=1e-8, smoother="chebyshev", pre=3, post=3):
        """saena_object::solve_smoother: `pre` sweeps of the smoother per iteration, nothing else"""
        return self._solve(lib().so_solve_smoother, rhs, max_iter, tol, smoother, pre, post)
