"""Kernel-mapping sweep on a large level-0 operator (tuning tool, not a test or the bench).
Builds the 7-point Poisson operator of experiments/Poisson.cpp's shape (interior unknowns only)
directly with numpy, wraps it in a dummy 2-level hierarchy, and times matvec / fused Chebyshev
sweep for every mapping."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from saena_b200.hierarchy import KIND_A, KIND_P, KIND_R, Hierarchy, Level, Operator  # noqa: E402
from saena_b200.native import Context  # noqa: E402


def poisson7(n):
    """n^3 unknowns, diag 6/h^2, off-diag -1/h^2 (aux_functions2.cpp:326-365), h = 1/(n+1)"""
    h2 = float((n + 1) ** 2)
    idx = np.arange(n ** 3, dtype=np.int64)
    i, j, k = idx % n, (idx // n) % n, idx // (n * n)
    cols = [idx - n * n, idx - n, idx - 1, idx, idx + 1, idx + n, idx + n * n]
    ok = [k > 0, j > 0, i > 0, np.ones_like(i, bool), i < n - 1, j < n - 1, k < n - 1]
    vals = [-h2, -h2, -h2, 6 * h2, -h2, -h2, -h2]
    C = np.stack(cols, 1)
    K = np.stack(ok, 1)
    V = np.broadcast_to(np.array(vals), C.shape)
    counts = K.sum(1).astype(np.int32)
    return counts, C[K].astype(np.int32), V[K].astype(np.float64)


def banded(n, b):
    idx = np.arange(n, dtype=np.int64)
    offs = np.arange(-b, b + 1)
    C = idx[:, None] + offs[None, :]
    K = (C >= 0) & (C < n)
    V = 1.0 / (idx[:, None] + C + 1.0)
    return K.sum(1).astype(np.int32), C[K].astype(np.int32), V[K]


def dummy_hierarchy(counts, cols, vals, n, nc=64):
    def op(kind, level, M, N, c, cc, v):
        return Operator(kind=kind, level=level, M=M, Mbig=M, Nbig=N, row_offset=0, col_offset=0, n_local_cols=N,
                        nnzPerRow_local=c, col_local=cc, val_local=v)
    pcol = (np.arange(n) % nc).astype(np.int32)
    order = np.argsort(pcol, kind="stable").astype(np.int32)
    rc = np.bincount(pcol, minlength=nc).astype(np.int32)
    eye = np.eye(nc)
    lv0 = Level(0, op(KIND_A, 0, n, n, counts, cols, vals), inv_diag=np.full(n, 1.0), eig_max=1.9,
                P=op(KIND_P, 0, n, nc, np.ones(n, np.int32), pcol, np.ones(n)),
                R=op(KIND_R, 0, nc, n, rc, order, np.ones(n)), M_coarse_old=nc, M_coarse=nc)
    lv1 = Level(1, op(KIND_A, 1, nc, nc, np.ones(nc, np.int32), np.arange(nc, dtype=np.int32), np.ones(nc)),
                inv_diag=np.ones(nc), eig_max=1.0)
    return Hierarchy([lv0, lv1], coarse_n=nc, coarse_row=np.arange(nc, dtype=np.int32),
                     coarse_col=np.arange(nc, dtype=np.int32), coarse_val=np.ones(nc))


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "poisson"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    t = time.time()
    if shape == "poisson":
        counts, cols, vals = poisson7(size)
        n = size ** 3
    else:
        band = int(sys.argv[3]) if len(sys.argv) > 3 else 64
        counts, cols, vals = banded(size, band)
        n = size
    h = dummy_hierarchy(counts, cols, vals, n)
    ctx = Context()
    ctx.upload_hierarchy(h)
    print(f"# {shape} n={n} nnz={len(vals)} built+uploaded in {time.time() - t:.1f}s", flush=True)
    mv_bytes = ctx.operator_bytes(0, KIND_A)
    sweep_bytes = 12 * len(vals) + 52 * n
    for m in [0, 100, 1, 2, 4, 8, 16, 32, -1, -2, -4]:
        ctx.set_mapping(0, KIND_A, m)
        ctx.time_matvec(0, KIND_A, 3)
        ms = ctx.time_matvec(0, KIND_A, 20)
        ms2 = ctx.time_smooth_sweep(0, "chebyshev", 20)
        print(json.dumps({"shape": shape, "n": n, "mapping": m, "matvec_ms": round(ms, 4),
                          "matvec_GBs": round(mv_bytes / ms / 1e6, 1), "cheb_first_sweep_ms": round(ms2, 4),
                          "cheb_GBs": round((sweep_bytes - 8 * n) / ms2 / 1e6, 1)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
