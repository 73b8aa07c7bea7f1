"""Smoothed-aggregation setup restated with tensor ops -- ONLY to feed the benchmark and the
large-size tests with a hierarchy of the reference's shape.

Why this exists.  The AMG setup is outside this package's scope: it stays on the reference's
host code (BASELINE.json north_star) and the drop-in adaptor uploads whatever
`saena_object::setup` produced.  But the reference's setup cannot feed the bench: measured in
this image it needs 3 s for 30^3 unknowns and 47 s for 64^3 (std::set assembly, recursive
SpGEMM), i.e. the better part of an hour for the 256^3 workload, and its sources do not exist on
the GPU box.  This module restates the setup's algorithm with torch tensor ops (GPU when there
is one) so that the bench solves on a hierarchy built by the SAME rules:

  strength of connection   create_strength_matrix   src/saena_object_setup1.cpp:520-720
                           + threshold              src/strength_matrix.cpp:233-270
  aggregation              aggregation_1_dist       src/saena_object_setup1.cpp:724-995
                           (a synchronous, data-parallel iteration: restated round for round,
                            so the aggregates are IDENTICAL to the reference's on one rank)
  coarse numbering         aggregate_index_update   src/saena_object_setup1.cpp:2103-2160
  prolongator              SA(): P = (I - w D^-1 A) P_t, w = float(2/3), drop |v| <= 1e-14
                                                    src/saena_object_setup1.cpp:8-254
  restriction              R = P^T                  src/restrict_matrix.cpp:10-494
  Galerkin product         Ac = R A P               src/saena_object_setup2.cpp:361
  filter                   small entries lumped to the diagonal, threshold x10 per level
                                                    src/saena_object_setup2.cpp:849-891
  level count              dynamic levels           src/saena_object_setup1.cpp:385-392
  Chebyshev bound          1.0001 * lambda_max(D^-1 A) by Lanczos, include/lamlan_saena.h:13-79
                           (the reference starts Lanczos from std::random_device and is not
                            reproducible run to run; here the start vector is seeded)
  precision flags          float_level              src/saena_object.cpp:241-244,277-285

tests/test_sa_setup.py checks the result against the reference's own hierarchy (same
aggregates / same operators to rounding) where oracle/_ref is available.  torch here is
plumbing for the setup only; nothing in the solve path imports this module.
"""
from __future__ import annotations

import math
import os
import sys
import time
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from .hierarchy import F64, I32, KIND_A, KIND_P, KIND_R, Hierarchy, Level, Operator

ALMOST_ZERO = 1e-14                  # include/data_struct.h:41
JACOBI_OMEGA = float(np.float32(2.0 / 3))   # saena_matrix.h:182 (a float, promoted)
BIG = 2 ** 62


@dataclass
class SetupOptions:
    """data/options006_poisson.xml"""
    conn_str: float = 0.2
    dynamic_levels: bool = True
    max_level: int = 20
    float_level: int = 0
    filter_thre: float = 1e-12
    filter_max: float = 1e-9
    filter_start: int = 1
    filter_rate: int = 1
    least_row_threshold: int = 100           # saena_object.h:43
    row_reduction_up_thrshld: float = 0.90   # saena_object.h:46
    lanczos_iters: int = 20
    seed: int = 2024


def _sync(dev):
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)


class _Csr:
    """row-major sorted COO/CSR triple on one torch device"""

    def __init__(self, n_rows, n_cols, row, col, val):
        self.n_rows, self.n_cols, self.row, self.col, self.val = n_rows, n_cols, row, col, val

    @property
    def nnz(self):
        return int(self.val.numel())

    def counts(self):
        return torch.bincount(self.row, minlength=self.n_rows)

    def diag(self):
        d = torch.zeros(self.n_rows, dtype=torch.float64, device=self.val.device)
        m = self.row == self.col
        d[self.row[m]] = self.val[m]
        return d


def _coalesce(n_rows, n_cols, row, col, val) -> _Csr:
    """sort by (row, col) and add duplicates"""
    key = row * n_cols + col
    key, order = torch.sort(key)
    val = val[order]
    uk, inv = torch.unique_consecutive(key, return_inverse=True)
    out = torch.zeros(uk.numel(), dtype=torch.float64, device=val.device)
    out.index_add_(0, inv, val)
    return _Csr(n_rows, n_cols, uk // n_cols, uk % n_cols, out)


# The one-process setup routes its products through the library's device SpGEMM since it passed
# tests/test_zzz_spgemm_gpu.py on a B200 (scipy, the reference's golden hierarchies, the live reference at 32^3, route
# equality at 64^3) and built the 256^3 bench hierarchy end to end (profiles/r02_setup_time.json,
# r02_bench_n1_native_spgemm.json).  The distributed setup keeps the tensor-op route until its own run with the device
# SpGEMM (SAENA_SETUP_SPGEMM=native asks for it there too).
SPGEMM_DEFAULT = "native"


def _native_spgemm_enabled(dev, dist: bool = False) -> bool:
    """SAENA_SETUP_SPGEMM=native: the library's own SpGEMM (csrc/spgemm.cu) does the products on a CUDA
    device; =torch: the tensor-op route below (the CPU path, and the A/B)."""
    return dev.type == "cuda" and os.environ.get("SAENA_SETUP_SPGEMM", "torch" if dist else SPGEMM_DEFAULT) != "torch"


def _spgemm_native(A: _Csr, B: _Csr) -> _Csr:
    """C = A B with the hand-written device kernels of libsaena_b200.so (saena_b200_spgemm_symbolic / _numeric)"""
    from . import native
    dev = A.val.device

    def csr_arrays(M: _Csr):
        rp = torch.zeros(M.n_rows + 1, dtype=torch.int64, device=dev)
        rp[1:] = torch.cumsum(M.counts(), 0)
        return rp, M.col.to(torch.int32).contiguous(), M.val.contiguous()

    a_rp, a_col, a_val = csr_arrays(A)
    b_rp, b_col, b_val = csr_arrays(B)
    c_rp, c_col, c_val = native.spgemm_csr(A.n_rows, A.n_cols, B.n_cols, a_rp, a_col, a_val, b_rp, b_col, b_val)
    row = torch.repeat_interleave(torch.arange(A.n_rows, device=dev), c_rp[1:] - c_rp[:-1])
    return _Csr(A.n_rows, B.n_cols, row, c_col.to(torch.int64), c_val)


def _spgemm(A: _Csr, B: _Csr, chunk_products: int = 300_000_000, dist: bool = False) -> _Csr:
    """C = A B: the library's device SpGEMM on CUDA; on the CPU (and for SAENA_SETUP_SPGEMM=torch) expand / sort /
    compress with tensor ops, in row chunks bounded by the number of products"""
    dev = A.val.device
    if _native_spgemm_enabled(dev, dist) and A.n_rows < 2 ** 31 and B.n_cols < 2 ** 31 - 1:
        return _spgemm_native(A, B)
    b_counts = B.counts()
    b_ptr = torch.zeros(B.n_rows + 1, dtype=torch.int64, device=dev)
    b_ptr[1:] = torch.cumsum(b_counts, 0)
    per_entry = b_counts[A.col]                       # products each A entry generates
    a_counts = A.counts()
    a_ptr = torch.zeros(A.n_rows + 1, dtype=torch.int64, device=dev)
    a_ptr[1:] = torch.cumsum(a_counts, 0)
    prod_cum = torch.zeros(A.nnz + 1, dtype=torch.int64, device=dev)
    prod_cum[1:] = torch.cumsum(per_entry, 0)
    rows_out, cols_out, vals_out = [], [], []
    r0 = 0
    row_prod_cum = prod_cum[a_ptr]                    # products before each row
    total = int(prod_cum[-1])
    while r0 < A.n_rows:
        # largest r1 with products(r0..r1) <= chunk
        target = int(row_prod_cum[r0]) + chunk_products
        r1 = int(torch.searchsorted(row_prod_cum, torch.tensor([target], device=dev), right=True)[0]) - 1
        r1 = max(r1, r0 + 1)
        r1 = min(r1, A.n_rows)
        e0, e1 = int(a_ptr[r0]), int(a_ptr[r1])
        if e1 > e0:
            pe = per_entry[e0:e1]
            n_prod = int(pe.sum())
            if n_prod:
                src = torch.repeat_interleave(torch.arange(e0, e1, device=dev), pe)      # A entry of each product
                first = prod_cum[e0:e1] - prod_cum[e0]
                within = torch.arange(n_prod, device=dev) - torch.repeat_interleave(first, pe)
                bidx = b_ptr[A.col[src]] + within
                c = _coalesce(A.n_rows, B.n_cols, A.row[src], B.col[bidx], A.val[src] * B.val[bidx])
                rows_out.append(c.row); cols_out.append(c.col); vals_out.append(c.val)
        r0 = r1
    del total
    if not rows_out:
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        return _Csr(A.n_rows, B.n_cols, z, z.clone(), torch.zeros(0, dtype=torch.float64, device=dev))
    return _Csr(A.n_rows, B.n_cols, torch.cat(rows_out), torch.cat(cols_out), torch.cat(vals_out))


def _to_torch_csr(A: _Csr) -> torch.Tensor:
    crow = torch.zeros(A.n_rows + 1, dtype=torch.int64, device=A.val.device)
    crow[1:] = torch.cumsum(A.counts(), 0)
    return torch.sparse_csr_tensor(crow, A.col, A.val, size=(A.n_rows, A.n_cols))


def _galerkin(R: _Csr, A: _Csr, P: _Csr, dense_budget_bytes: float = 48e9) -> _Csr:
    """Ac = R A P.  On the deep levels of a 3D hierarchy the product is nearly dense (256^3 Poisson,
    level 3 -> 4: rows of A P have ~20 000 entries, 8e11 scalar products) and the sparse expansion
    drowns; there the two products are done as sparse x dense (R (A P_dense)) and the result is
    re-sparsified.  Structural zeros stay exact zeros, so the pattern is the sparse product's
    (barring exact cancellation)."""
    nc = P.n_cols
    dense_bytes = 8.0 * A.n_rows * nc
    avg_prod_per_row = (A.nnz / max(A.n_rows, 1)) * (P.nnz / max(P.n_rows, 1))
    if A.val.is_cuda and not _native_spgemm_enabled(A.val.device) and 2 * dense_bytes <= dense_budget_bytes \
            and avg_prod_per_row > 4 * nc:
        Pd = torch.zeros(P.n_rows, nc, dtype=torch.float64, device=A.val.device)
        Pd[P.row, P.col] = P.val
        AP = torch.sparse.mm(_to_torch_csr(A), Pd)
        del Pd
        Ac = torch.sparse.mm(_to_torch_csr(R), AP)
        del AP
        idx = torch.nonzero(Ac, as_tuple=True)
        return _Csr(nc, nc, idx[0], idx[1], Ac[idx])
    return _spgemm(R, _spgemm(A, P))


def _transpose(A: _Csr) -> _Csr:
    key = A.col * A.n_rows + A.row
    key, order = torch.sort(key)
    return _Csr(A.n_cols, A.n_rows, key // A.n_rows, key % A.n_rows, A.val[order])


def _matvec(A: _Csr, x: torch.Tensor) -> torch.Tensor:
    y = torch.zeros(A.n_rows, dtype=torch.float64, device=x.device)
    y.index_add_(0, A.row, A.val * x[A.col])
    return y


# ------------------------------------------------------------------------------------------
def strength_graph(A: _Csr, conn_str: float):
    """strong off-diagonal connections (row, col) -- create_strength_matrix + setup_matrix"""
    n, dev = A.n_rows, A.val.device
    off = A.row != A.col
    max_per_row = torch.full((n,), -torch.finfo(torch.float64).max, dtype=torch.float64, device=dev)
    max_per_row.scatter_reduce_(0, A.row[off], -A.val[off], reduce="amax")
    s_row = -A.val / max_per_row[A.row]      # entry   (normalised by its row's maximum)
    s_col = -A.val / max_per_row[A.col]      # entryT  (normalised by the column index's row maximum)
    strong = off & ((s_row > conn_str) | (s_col > conn_str))
    return A.row[strong], A.col[strong]


def _expand_segments(ptr: torch.Tensor, rows: torch.Tensor):
    """entry indices of the CSR segments of `rows` and, per entry, the position of its row in `rows`"""
    dev = rows.device
    start = ptr[rows]
    deg = ptr[rows + 1] - start
    total = int(deg.sum())
    seg = torch.repeat_interleave(torch.arange(rows.numel(), device=dev), deg, output_size=total)
    first = torch.cumsum(deg, 0) - deg
    e = start[seg] + (torch.arange(total, device=dev) - first[seg])
    return e, seg


def aggregate(n: int, s_row: torch.Tensor, s_col: torch.Tensor):
    """aggregation_1_dist on one rank.  The reference sweeps every undecided node in every round;
    a node's outcome can only change when one of its neighbours was decided in the round before,
    so only those nodes are re-evaluated here (same rule, same synchronous updates, hence the same
    aggregates -- tests/test_sa_setup.py compares with the reference's).  Returns (aggregate id per
    node in coarse numbering, number of aggregates, rounds)."""
    dev = s_row.device
    # CSR of the strength graph by row (s_row is ascending: A is row-major sorted) and by column
    ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    ptr[1:] = torch.cumsum(torch.bincount(s_row, minlength=n), 0)
    order = torch.argsort(s_col, stable=True)
    t_nb = s_row[order]                      # rows that list node j as a neighbour, grouped by j
    tptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    tptr[1:] = torch.cumsum(torch.bincount(s_col, minlength=n), 0)
    del order
    agg = torch.arange(n, dtype=torch.int64, device=dev)
    decided = torch.zeros(n, dtype=torch.bool, device=dev)
    is_root = torch.zeros(n, dtype=torch.bool, device=dev)
    F = torch.arange(n, dtype=torch.int64, device=dev)   # undecided nodes to evaluate this round
    rounds = 0
    while F.numel():
        rounds += 1
        e, seg = _expand_segments(ptr, F)
        nb = s_col[e]
        # eligible neighbours: undecided or root; their aggregate value is their own index
        elig = (~decided[nb]) | is_root[nb]
        cand = torch.where(elig, nb, torch.full_like(nb, BIG))
        m = torch.full((F.numel(),), BIG, dtype=torch.int64, device=dev)
        m.scatter_reduce_(0, seg, cand, reduce="amin")
        take = m < F                                     # a smaller eligible neighbour exists
        mc = torch.where(take, m, F)
        dec_nei = torch.where(take, decided[mc], torch.ones_like(take))
        root_nei = take & is_root[mc]
        new_root = dec_nei & (~take)
        join = dec_nei & root_nei
        newly = F[dec_nei]
        # synchronous update: everything above read the state of the previous round
        agg[F[join]] = mc[join]
        is_root[F[new_root]] = True
        decided[newly] = True
        # next round: undecided nodes one of whose neighbours was just decided
        te, _ = _expand_segments(tptr, newly)
        cand_rows = t_nb[te]
        cand_rows = cand_rows[~decided[cand_rows]]
        F = torch.unique(cand_rows)
    assert bool(decided.all())
    # aggregate_index_update: coarse id = rank of the root among the sorted roots
    coarse_of_root = torch.cumsum(is_root.to(torch.int64), 0) - 1
    return coarse_of_root[agg], int(is_root.sum()), rounds


def prolongator(A: _Csr, agg_c: torch.Tensor, nc: int, inv_diag: torch.Tensor) -> _Csr:
    """SA(): P = (I - w D^-1 A) P_t with P_t(i, agg(i)) = 1"""
    v = -JACOBI_OMEGA * inv_diag[A.row] * A.val
    v = torch.where(A.row == A.col, v + 1.0, v)
    P = _coalesce(A.n_rows, nc, A.row, agg_c[A.col], v)
    keep = P.val.abs() > ALMOST_ZERO
    return _Csr(P.n_rows, nc, P.row[keep], P.col[keep], P.val[keep])


def filter_entries(Ac: _Csr, thre: float) -> _Csr:
    """saena_object::filter: off-diagonal |v| <= thre is lumped into the diagonal"""
    n, dev = Ac.n_rows, Ac.val.device
    is_diag = Ac.row == Ac.col
    keep = (Ac.val.abs() > thre) | is_diag
    add = torch.zeros(n, dtype=torch.float64, device=dev)
    add.index_add_(0, Ac.row[~keep], Ac.val[~keep])
    row, col, val = Ac.row[keep], Ac.col[keep], Ac.val[keep].clone()
    d = row == col
    val[d] = val[d] + add[row[d]]
    val[d] = torch.where(val[d].abs() < ALMOST_ZERO, torch.ones_like(val[d]), val[d])
    has_diag = torch.zeros(n, dtype=torch.bool, device=dev)
    has_diag[row[d]] = True
    if not bool(has_diag.all()):
        miss = torch.nonzero(~has_diag).flatten()
        return _coalesce(n, n, torch.cat([row, miss]), torch.cat([col, miss]),
                         torch.cat([val, torch.ones(miss.numel(), dtype=torch.float64, device=dev)]))
    return _Csr(n, n, row, col, val)


def lanczos_eig_max(A: _Csr, inv_diag: torch.Tensor, iters: int, seed: int) -> float:
    """largest eigenvalue of D^-1/2 A D^-1/2 (find_eig: scale_matrix + find_eig_lamlan), x 1.0001"""
    n, dev = A.n_rows, A.val.device
    s = torch.sqrt(inv_diag.abs())
    g = torch.Generator(device="cpu").manual_seed(seed)
    v = torch.randn(n, dtype=torch.float64, generator=g).to(dev)
    v /= torch.linalg.norm(v)
    v_prev = torch.zeros_like(v)
    alphas, betas = [], []
    beta = 0.0
    k = min(iters, n)
    for _ in range(k):
        w = s * _matvec(A, s * v) - beta * v_prev
        alpha = float(torch.dot(w, v))
        w = w - alpha * v
        beta = float(torch.linalg.norm(w))
        alphas.append(alpha)
        if beta < 1e-14:
            break
        betas.append(beta)
        v_prev, v = v, w / beta
    T = np.diag(alphas) + np.diag(betas[:len(alphas) - 1], 1) + np.diag(betas[:len(alphas) - 1], -1)
    return 1.0001 * float(np.linalg.eigvalsh(T)[-1])


# ------------------------------------------------------------------------------------------
@dataclass
class DeviceLevel:
    A: _Csr
    inv_diag: torch.Tensor
    eig_max: float
    a_use_double: bool
    P: Optional[_Csr] = None
    R: Optional[_Csr] = None
    pr_use_double: bool = True


class DeviceHierarchy:
    """The global (one-rank) hierarchy as torch tensors, plus its conversion to any rank's share
    of a 1-D row partition in the reference layout (`saena_b200.hierarchy.Operator`)."""

    def __init__(self, levels: List[DeviceLevel]):
        self.levels = levels

    def summary(self) -> str:
        return "\n".join(f"L{l}: M={lv.A.n_rows} nnz={lv.A.nnz} ({lv.A.nnz / lv.A.n_rows:.1f}/row) eig={lv.eig_max:.4f}"
                         + (f" | P nnz={lv.P.nnz}" if lv.P is not None else "") for l, lv in enumerate(self.levels))

    # -- partitions ----------------------------------------------------------------------------
    @staticmethod
    def _balanced_split(M: _Csr, nprocs: int) -> np.ndarray:
        """nnz-balanced contiguous row blocks (hierarchy.balanced_split on device counts)"""
        from .hierarchy import balanced_split, csr_from_counts
        return balanced_split(csr_from_counts(M.counts().cpu().numpy()), nprocs)

    @staticmethod
    def _aligned_coarse_split(R: _Csr, fine_split: np.ndarray) -> np.ndarray:
        """hierarchy.aligned_coarse_split on device tensors"""
        dev = R.val.device
        nprocs = len(fine_split) - 1
        nc = R.n_rows
        a = R.val.abs()
        row_max = torch.zeros(nc, dtype=torch.float64, device=dev)
        row_max.scatter_reduce_(0, R.row, a, reduce="amax", include_self=False)
        cand = torch.where(a == row_max[R.row], R.col, torch.full_like(R.col, BIG))
        arg_col = torch.full((nc,), BIG, dtype=torch.int64, device=dev)
        arg_col.scatter_reduce_(0, R.row, cand, reduce="amin")
        fs = torch.as_tensor(np.asarray(fine_split, np.int64), device=dev)
        owner = torch.clamp(torch.searchsorted(fs, arg_col, right=True) - 1, 0, nprocs - 1)
        owner = torch.cummax(owner, 0).values
        return torch.searchsorted(owner, torch.arange(nprocs + 1, device=dev), right=False).cpu().numpy().astype(np.int64)

    def splits(self, nprocs: int, agglomerate_below: int, rebalance_above: float = 0.0):
        """(row partition of every level, partition R of the level above writes into, agglomerated?)
        -- hierarchy.partition_hierarchy's rules"""
        L = len(self.levels)
        splits, aligned, agglomerated = [], [], []
        for l, lv in enumerate(self.levels):
            agg = l > 0 and (lv.A.n_rows < agglomerate_below or agglomerated[-1] or l == L - 1) and nprocs > 1
            agglomerated.append(bool(agg))
            all_on_0 = np.concatenate(([0], np.full(nprocs, lv.A.n_rows))).astype(np.int64)
            if l == 0:
                aligned.append(None)
                splits.append(self._balanced_split(lv.A, nprocs).astype(np.int64))
                continue
            al = all_on_0 if agglomerated[l - 1] else self._aligned_coarse_split(self.levels[l - 1].R, splits[l - 1])
            if rebalance_above > 0 and not agg and not agglomerated[l - 1]:
                from .hierarchy import csr_from_counts, split_imbalance
                indptr = csr_from_counts(lv.A.counts().cpu().numpy())
                if split_imbalance(indptr, al) > rebalance_above:
                    al = self._balanced_split(lv.A, nprocs).astype(np.int64)
            aligned.append(al)
            splits.append(all_on_0 if agg else al)
        return splits, aligned, agglomerated

    @staticmethod
    def _rank_operator(kind, level, M: _Csr, row_split, col_split, rank, use_double) -> Operator:
        """one rank's Operator -- hierarchy.split_operator restated on device tensors, building
        only `rank`'s arrays (the other ranks' requests are still needed for the send plan)"""
        dev = M.val.device
        nprocs = len(row_split) - 1
        rs = torch.as_tensor(row_split, device=dev)
        cs = torch.as_tensor(col_split, device=dev)
        r0, r1, c0, c1 = int(row_split[rank]), int(row_split[rank + 1]), int(col_split[rank]), int(col_split[rank + 1])
        mine = (M.row >= r0) & (M.row < r1)
        row, col, val = M.row[mine] - r0, M.col[mine], M.val[mine]
        local = (col >= c0) & (col < c1)
        nrows = r1 - r0
        op = dict(kind=kind, level=level, M=nrows, Mbig=M.n_rows, Nbig=M.n_cols, row_offset=r0, col_offset=c0,
                  n_local_cols=c1 - c0, use_double=use_double, nprocs=nprocs, rank=rank)
        op["nnzPerRow_local"] = torch.bincount(row[local], minlength=nrows).to(torch.int32).cpu().numpy()
        op["col_local"] = col[local].to(torch.int32).cpu().numpy()
        op["val_local"] = val[local].cpu().numpy()
        if nprocs == 1:
            return Operator(**op)
        # remote part: column-major (col, then row)
        rr, rc, rv = row[~local], col[~local], val[~local]
        order = torch.argsort(rc * max(nrows, 1) + rr)
        rr, rc, rv = rr[order], rc[order], rv[order]
        distinct, counts = torch.unique_consecutive(rc, return_counts=True)
        owner_d = torch.searchsorted(cs, distinct, right=True) - 1
        owner_e = torch.searchsorted(cs, rc, right=True) - 1
        recv_count = torch.bincount(owner_d, minlength=nprocs).cpu().numpy().astype(np.int32)
        nnz_per_proc = torch.bincount(owner_e, minlength=nprocs).cpu().numpy()
        op["row_remote"] = rr.to(torch.int32).cpu().numpy()
        op["val_remote"] = rv.cpu().numpy()
        op["nnzPerCol_remote"] = counts.to(torch.int32).cpu().numpy()
        op["nnzPerProcScan"] = np.concatenate(([0], np.cumsum(nnz_per_proc))).astype(np.int64)
        op["rdispls"] = np.concatenate(([0], np.cumsum(recv_count)[:-1])).astype(np.int32)
        op["recvProcRank"] = np.flatnonzero(recv_count).astype(np.int32)
        op["recvProcCount"] = recv_count[recv_count != 0]
        # what the others ask of me: entries whose column I own and whose row I do not
        wanted = (M.col >= c0) & (M.col < c1) & ~mine
        dst = torch.searchsorted(rs, M.row[wanted], right=True) - 1
        key = torch.unique(dst * M.n_cols + M.col[wanted])
        dst_u, col_u = key // M.n_cols, key % M.n_cols
        send_count = torch.bincount(dst_u, minlength=nprocs).cpu().numpy().astype(np.int32)
        op["vIndex"] = (col_u - c0).to(torch.int32).cpu().numpy()
        op["vdispls"] = np.concatenate(([0], np.cumsum(send_count)[:-1])).astype(np.int32)
        op["sendProcRank"] = np.flatnonzero(send_count).astype(np.int32)
        op["sendProcCount"] = send_count[send_count != 0]
        return Operator(**op)

    def to_rank(self, rank: int = 0, nprocs: int = 1, agglomerate_below: int = 0,
                rebalance_above: float = 0.0) -> Hierarchy:
        """This rank's share (hierarchy.partition_hierarchy semantics: nnz-balanced row blocks per
        level, levels below `agglomerate_below` global rows -- and always the coarsest -- on rank 0)."""
        splits, aligned, agglomerated = self.splits(nprocs, agglomerate_below, rebalance_above)
        levels = []
        for l, lv in enumerate(self.levels):
            sp = splits[l]
            r0, r1 = int(sp[rank]), int(sp[rank + 1])
            A = self._rank_operator(KIND_A, l, lv.A, sp, sp, rank, lv.a_use_double)
            out = Level(level=l, A=A, inv_diag=lv.inv_diag[r0:r1].cpu().numpy(), eig_max=lv.eig_max, active=r1 > r0)
            if lv.P is not None:
                so, sn = aligned[l + 1], splits[l + 1]
                out.P = self._rank_operator(KIND_P, l, lv.P, sp, so, rank, lv.pr_use_double)
                out.R = self._rank_operator(KIND_R, l, lv.R, so, sp, rank, lv.pr_use_double)
                out.M_coarse_old = int(so[rank + 1] - so[rank])
                out.M_coarse = int(sn[rank + 1] - sn[rank])
                if not np.array_equal(so, sn):
                    for peer in range(nprocs):
                        a, b = max(so[rank], sn[peer]), min(so[rank + 1], sn[peer + 1])
                        if b > a:
                            out.repart_send.append((peer, int(a - so[rank]), int(b - a)))
                        a, b = max(sn[rank], so[peer]), min(sn[rank + 1], so[peer + 1])
                        if b > a:
                            out.repart_recv.append((peer, int(a - sn[rank]), int(b - a)))
            levels.append(out)
        last = self.levels[-1].A
        return Hierarchy(levels=levels, coarse_n=last.n_rows, coarse_row=last.row.to(torch.int32).cpu().numpy(),
                         coarse_col=last.col.to(torch.int32).cpu().numpy(), coarse_val=last.val.cpu().numpy(),
                         nprocs=nprocs, rank=rank)


def build_device_hierarchy(n: int, row, col, val, opts: Optional[SetupOptions] = None, device: Optional[str] = None,
                           verbose: bool = False) -> DeviceHierarchy:
    """Global hierarchy from a square matrix given as COO (any order, duplicates added)."""
    opts = opts or SetupOptions()
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))

    def as_t(a, dt):
        return a.to(dev, dt) if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dt)

    A = _coalesce(n, n, as_t(row, torch.int64), as_t(col, torch.int64), as_t(val, torch.float64))
    levels: List[DeviceLevel] = []
    filter_thre = opts.filter_thre
    filter_it = 0
    l = 0
    last_level = False
    a_use_double = opts.float_level != 0                  # saena_object.cpp:241-244
    while True:
        d = A.diag()
        if bool((d.abs() < ALMOST_ZERO).any()):
            raise ValueError("zero diagonal element")      # inverse_diag: print + exit in the reference
        inv_diag = 1.0 / d
        eig = lanczos_eig_max(A, inv_diag, opts.lanczos_iters, opts.seed + l)
        lv = DeviceLevel(A=A, inv_diag=inv_diag, eig_max=eig, a_use_double=a_use_double)
        levels.append(lv)
        if verbose:
            print(f"level {l}: rows {A.n_rows} nnz {A.nnz} ({A.nnz / A.n_rows:.1f}/row) eig {eig:.4f}", file=sys.stderr, flush=True)
        if l >= opts.max_level or last_level:
            break
        t_ag = time.perf_counter()
        s_r, s_c = strength_graph(A, opts.conn_str)
        agg_c, nc, rounds = aggregate(A.n_rows, s_r, s_c)
        del s_r, s_c
        if verbose:
            _sync(dev)
            print(f"   aggregation: {nc} aggregates in {rounds} rounds, {time.perf_counter() - t_ag:.1f}s", file=sys.stderr, flush=True)
        t_rap = time.perf_counter()
        # find_aggregation's dynamic-level rule: the level about to be created is the last one
        if opts.dynamic_levels:
            last_level = bool(nc <= opts.least_row_threshold or
                              np.float32(nc) / np.float32(A.n_rows) > np.float32(opts.row_reduction_up_thrshld))
        else:
            last_level = (l + 1 == opts.max_level)
        P = prolongator(A, agg_c, nc, inv_diag)
        R = _transpose(P)
        Ac = _galerkin(R, A, P)
        filter_it += 1
        if filter_it >= opts.filter_start:
            filter_thre = min(filter_thre, opts.filter_max)
            Ac = filter_entries(Ac, filter_thre)
            filter_thre *= 10 ** opts.filter_rate
        if verbose:
            _sync(dev)
            print(f"   P, R, RAP, filter: {time.perf_counter() - t_rap:.1f}s", file=sys.stderr, flush=True)
        lv.P, lv.R = P, R
        lv.pr_use_double = not (l >= opts.float_level)     # saena_object.cpp:277-280
        a_use_double = not (l + 1 >= opts.float_level)     # :281-284
        A = Ac
        l += 1
    return DeviceHierarchy(levels)


def build_hierarchy(n: int, row, col, val, opts: Optional[SetupOptions] = None, device: Optional[str] = None,
                    verbose: bool = False) -> Hierarchy:
    """One-rank hierarchy in the reference layout."""
    return build_device_hierarchy(n, row, col, val, opts, device, verbose).to_rank(0, 1)


# ------------------------------------------------------------------------------------------
# the synthetic inputs of SURVEY.md 8d
# ------------------------------------------------------------------------------------------
def poisson3d_coo(n: int):
    """7-point Laplacian on the n^3 INTERIOR nodes of an (n+2)^3 grid: what laplacian3D(mx = n+2) +
    set_remove_boundary(true) leaves (src/aux_functions2.cpp:326-365): diag 6/h^2, off-diag -1/h^2,
    h = 1/(mx-1)."""
    mx = n + 2
    h2 = float((mx - 1) ** 2)
    idx = np.arange(n ** 3, dtype=np.int64)
    i, j, k = idx % n, (idx // n) % n, idx // (n * n)
    rows, cols, vals = [idx], [idx], [np.full(idx.shape, 6.0 * h2)]
    for ok, off in ((i > 0, -1), (i < n - 1, 1), (j > 0, -n), (j < n - 1, n), (k > 0, -n * n), (k < n - 1, n * n)):
        rows.append(idx[ok]); cols.append(idx[ok] + off); vals.append(np.full(int(ok.sum()), -h2))
    return n ** 3, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)


def poisson3d_rhs(n: int) -> np.ndarray:
    """laplacian3D_set_rhs restricted to interior nodes (src/aux_functions2.cpp:629-700):
    12 pi^2 sin(2 pi x) sin(2 pi y) sin(2 pi z) at x = i/(mx-1)"""
    mx = n + 2
    t = np.arange(1, mx - 1, dtype=np.float64) / (mx - 1)
    s = np.sin(2 * math.pi * t)
    return (12 * math.pi ** 2 * s[None, None, :] * s[None, :, None] * s[:, None, None]).ravel()


# forward half of a 2-D neighbourhood ordered by distance: node (x, y) couples to (x+dx, y+dy); the
# mirrored entries make the pattern symmetric
_HALF_OFFSETS = [(1, 0), (0, 1), (1, 1), (-1, 1), (2, 0), (0, 2), (2, 1), (1, 2), (-1, 2), (-2, 1), (2, 2), (-2, 2),
                 (3, 0), (0, 3), (3, 1), (1, 3), (-1, 3), (-3, 1), (3, 2), (2, 3), (-2, 3), (-3, 2), (4, 0), (0, 4)]


def unstructured2d_coo(g: int, seed: int = 2024, row_lengths=(6, 8, 10), weights=(16, 48, 20), tile: int = 16,
                       shift: float = 0.05):
    """Synthetic 2-D Helmholtz-like matrix with an unstructured-mesh pattern (BASELINE.json configs[4],
    SURVEY.md 8d): g*g nodes; the number of entries per row is drawn from the empirical distribution of
    the shape donors data/Helmholtz2D_CG_curved_tri/*.mtx (P2: 6/8/10 entries in proportion 16:48:20;
    pass row_lengths=(24, 32, 40) for the P8 shape); every node couples to its nearest neighbours until
    its drawn length is reached, the pattern is then symmetrised; nodes are numbered tile by tile
    with a random order inside each `tile` x `tile` patch (mesh-generator-like locality, irregular
    gathers); off-diagonals -(0.5 + u), u uniform, symmetric; diagonal = sum |off-diagonal| * (1 + shift):
    SPD, the positive shift being the Helmholtz term.  Returns (n, row, col, val) with duplicates-free COO."""
    rng = np.random.default_rng(seed)
    n = g * g
    x, y = np.meshgrid(np.arange(g), np.arange(g), indexing="xy")
    x, y = x.ravel(), y.ravel()
    # numbering: tiles in row-major order, random order inside a tile
    tx, ty = x // tile, y // tile
    key = (ty * ((g + tile - 1) // tile) + tx).astype(np.int64) * (tile * tile * 4) + rng.permutation(n) % (tile * tile * 4)
    order = np.argsort(key, kind="stable")
    number = np.empty(n, np.int64)
    number[order] = np.arange(n)
    # forward edges until the drawn row length is reached (half of the off-diagonal count each way)
    want = rng.choice(np.asarray(row_lengths), size=n, p=np.asarray(weights, float) / np.sum(weights))
    half = (want - 1 + 1) // 2
    src, dst = [], []
    for k, (dx, dy) in enumerate(_HALF_OFFSETS[:int(half.max())]):
        ok = (half > k) & (x + dx >= 0) & (x + dx < g) & (y + dy < g)
        src.append(np.flatnonzero(ok))
        dst.append(src[-1] + dx + dy * g)
    src, dst = np.concatenate(src), np.concatenate(dst)
    w = -(0.5 + rng.uniform(0, 1, len(src)))
    i, j = number[src], number[dst]
    diag = np.zeros(n)
    np.add.at(diag, i, -w)
    np.add.at(diag, j, -w)
    diag *= 1.0 + shift
    diag[diag == 0] = 1.0
    row = np.concatenate((i, j, np.arange(n)))
    col = np.concatenate((j, i, np.arange(n)))
    val = np.concatenate((w, w, diag))
    return n, row, col, val


def unstructured2d_rhs(n: int, seed: int = 2024) -> np.ndarray:
    return np.random.default_rng(seed + 1).uniform(-1, 1, n)
