"""Run under torchrun with 2 CPU processes (gloo).  Exercises the host side of the N>1 path:
every rank partitions the same one-rank hierarchy, keeps its share, exchanges ghost values with
its peers following the uploaded plan (vIndex / vdispls / rdispls, as the NCCL path does with
ncclSend/ncclRecv), applies local + remote parts on the CPU with numpy and compares with the
oracle's emulation of the same partition."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.oracle import Oracle  # noqa: E402
from saena_b200.hierarchy import partition_hierarchy  # noqa: E402
from saena_b200.distributed import all_ranks_ok, exchange_nccl_id, halo_exchange_host  # noqa: E402
from tests.util import GOLDEN, Golden  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = Golden(GOLDEN[1])
    hs = partition_hierarchy(g.hier, world, agglomerate_below=100)
    mine = hs[rank]
    # the id broadcast used to bootstrap NCCL works over any backend
    ident = exchange_nccl_id(lambda: bytes(range(128)))
    assert ident == bytes(range(128))
    # the agreement every fallback decision of bench.py rests on: a failure on ANY rank is a failure on every rank
    assert all_ranks_ok(True) is True
    assert all_ranks_ok(rank != world - 1) is False
    assert all_ranks_ok(rank == world - 1) is (world == 1)
    rng = np.random.default_rng(5)
    for l, lv in enumerate(mine.levels):
        A = lv.A
        full = rng.standard_normal(A.Mbig)
        v = full[A.row_offset:A.row_offset + A.M]
        ghost = halo_exchange_host(A, v)
        w = A.to_scipy_local() @ v
        k = 0
        for j, cnt in enumerate(A.nnzPerCol_remote):
            for _ in range(cnt):
                w[A.row_remote[k]] += A.val_remote[k] * ghost[j]
                k += 1
        sizes = [h.levels[l].A.M for h in hs]
        off = np.concatenate(([0], np.cumsum(sizes)))
        ref = Oracle(hs).matvec(l, 0, [full[off[i]:off[i + 1]] for i in range(world)])[rank]
        err = np.linalg.norm(w - ref) / max(np.linalg.norm(ref), 1e-300)
        assert err < 1e-13, (l, err)
    dist.barrier()
    if rank == 0:
        print("HALO_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
