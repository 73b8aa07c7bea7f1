"""Random operators and random row / column partitions (empty blocks included) through hierarchy.split_operator:
the per-rank reference layout must carry exactly the matrix (local block + remote block re-assembled through the
senders' vIndex lists) and the multi-rank oracle's product on it must be the global product.  Seeded sweep (no
shrinking needed: every failing case prints its seed)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle.oracle import Oracle
from saena_b200.hierarchy import KIND_A, KIND_R, Hierarchy, Level, split_operator


def _random_split(rng, n, nprocs):
    cuts = np.sort(rng.integers(0, n + 1, nprocs - 1))
    if rng.uniform() < 0.3 and nprocs > 2:      # force an empty block
        cuts[1] = cuts[0]
    return np.concatenate(([0], cuts, [n])).astype(np.int64)


@pytest.mark.parametrize("seed", range(12))
def test_split_operator_carries_the_matrix_and_the_oracle_multiplies_it(seed):
    rng = np.random.default_rng(seed)
    nprocs = int(rng.integers(1, 6))
    n_rows, n_cols = int(rng.integers(1, 60)), int(rng.integers(1, 60))
    square = rng.uniform() < 0.5
    if square:
        n_cols = n_rows
    A = sp.random(n_rows, n_cols, density=float(rng.uniform(0.02, 0.5)), format="csr", random_state=seed, dtype=np.float64)
    A.sort_indices()
    rs = _random_split(rng, n_rows, nprocs)
    cs = rs if square else _random_split(rng, n_cols, nprocs)
    use_double = bool(rng.uniform() < 0.5)
    ops = split_operator(KIND_A if square else KIND_R, 0, A.indptr, A.indices, A.data, n_cols, rs, cs, use_double)
    # 1. the layout carries the matrix
    rows, cols, vals = [], [], []
    for op in ops:
        r = np.repeat(np.arange(op.M), op.nnzPerRow_local) + op.row_offset
        rows.append(r); cols.append(op.col_local.astype(np.int64)); vals.append(op.val_local)
        ghost = []
        for p, cnt in zip(op.recvProcRank, op.recvProcCount):
            peer = ops[int(p)]
            o = int(peer.vdispls[op.rank])
            ghost.append(peer.vIndex[o:o + int(cnt)].astype(np.int64) + peer.col_offset)
        ghost = np.concatenate(ghost) if ghost else np.zeros(0, np.int64)
        assert len(ghost) == op.col_remote_size
        rows.append(op.row_remote.astype(np.int64) + op.row_offset)
        cols.append(np.repeat(ghost, op.nnzPerCol_remote)); vals.append(op.val_remote)
    B = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=A.shape)
    assert (abs(A - B)).sum() == 0, seed
    # 2. the oracle's distributed product on it
    x = rng.uniform(-1, 1, n_cols)
    hs = [Hierarchy(levels=[Level(level=0, A=op, inv_diag=np.ones(op.M), eig_max=1.0)], nprocs=nprocs, rank=r)
          for r, op in enumerate(ops)]
    if square:
        y = Oracle(hs).matvec(0, KIND_A, [x[cs[r]:cs[r + 1]] for r in range(nprocs)])
        want = A @ x
        if not use_double and nprocs > 1:
            # ghost values travel as float: only the entries outside the diagonal blocks see the cast
            assert np.allclose(np.concatenate(y), want, rtol=0, atol=2e-7 * np.abs(A).sum(axis=1).max())
        else:
            assert np.allclose(np.concatenate(y), want, rtol=1e-13, atol=1e-14), seed
