"""Run under torchrun (gloo): sa_setup_dist.rank_operator -- each rank builds ITS operator from ITS rows, the send
side arriving through the request lists -- against hierarchy.split_operator, which builds every rank's operator
from the global matrix: every array of the reference layout must be identical.  Random rectangular operators,
random row and column partitions with empty blocks."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from saena_b200 import sa_setup_dist as sd  # noqa: E402
from saena_b200.hierarchy import _OP_ARRAYS, KIND_R, balanced_split, split_operator  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = sd.Comm()
    for seed in range(10):
        rng = np.random.default_rng(seed)
        n_rows, n_cols = int(rng.integers(1, 80)), int(rng.integers(1, 80))
        A = sp.random(n_rows, n_cols, density=float(rng.uniform(0.02, 0.4)), format="csr", random_state=seed, dtype=np.float64)
        A.sort_indices()

        def split(n):
            cuts = np.sort(rng.integers(0, n + 1, world - 1))
            if world > 2 and rng.uniform() < 0.4:
                cuts[1] = cuts[0]
            return np.concatenate(([0], cuts, [n])).astype(np.int64)

        rs, cs = split(n_rows), split(n_cols)
        want = split_operator(KIND_R, 3, A.indptr, A.indices, A.data, n_cols, rs, cs, use_double=False)[rank]
        r0, r1 = int(rs[rank]), int(rs[rank + 1])
        blk = A[r0:r1].tocoo()
        order = np.lexsort((blk.col, blk.row))
        M = sd.DCsr(n_rows, n_cols, rs, rank, torch.as_tensor(blk.row[order].astype(np.int64)),
                    torch.as_tensor(blk.col[order].astype(np.int64)), torch.as_tensor(blk.data[order]))
        got = sd.rank_operator(comm, KIND_R, 3, M, cs, use_double=False)
        for f in ("kind", "level", "M", "Mbig", "Nbig", "row_offset", "col_offset", "n_local_cols", "use_double", "nprocs", "rank"):
            assert getattr(got, f) == getattr(want, f), (seed, f, getattr(got, f), getattr(want, f))
        for f in _OP_ARRAYS:
            a, b = np.asarray(getattr(got, f)), np.asarray(getattr(want, f))
            assert a.shape == b.shape and np.array_equal(a, b), (seed, rank, f, a, b)
        # the nnz-balanced split found collectively = hierarchy.balanced_split on the global row offsets
        assert np.array_equal(sd.balanced_split_dist(comm, M), balanced_split(A.indptr.astype(np.int64), world)), seed
        # moving the rows to another partition and back changes nothing
        other = split(n_rows)
        back = sd.repartition(comm, sd.repartition(comm, M, other), rs)
        assert torch.equal(back.row, M.row) and torch.equal(back.col, M.col) and torch.equal(back.val, M.val), seed
        moved = sd.repartition(comm, M, other)
        o0, o1 = int(other[rank]), int(other[rank + 1])
        ref_blk = A[o0:o1].tocoo()
        ordr = np.lexsort((ref_blk.col, ref_blk.row))
        assert np.array_equal(moved.row.numpy(), ref_blk.row[ordr]) and np.array_equal(moved.col.numpy(), ref_blk.col[ordr]), seed
    dist.barrier()
    if rank == 0:
        print("RANK_OPERATOR_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
