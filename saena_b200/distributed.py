"""Process-group plumbing for the N>1 path: one process per GPU, launched by torchrun.

torch.distributed is used only to bootstrap (sharing the NCCL unique id, barriers, max-over-ranks
timing); the data path -- halo exchange, dot-product all-reduce, coarse-vector repartition -- is
NCCL called from libsaena_b200.so on its own streams.
"""
from __future__ import annotations

from typing import Callable

import numpy as np

from .hierarchy import Operator


def exchange_nccl_id(make_id: Callable[[], bytes]) -> bytes:
    """Rank 0 creates the id (saena_b200_nccl_unique_id), everyone receives the 128 bytes."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if rank == 0:
        buf = torch.tensor(list(make_id()), dtype=torch.uint8, device=dev)
    else:
        buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().tolist())


def halo_exchange_host(op: Operator, v: np.ndarray) -> np.ndarray:
    """The halo exchange of one operator with torch.distributed point-to-point on HOST buffers,
    following the same plan the device path executes with ncclSend/ncclRecv (pack with vIndex,
    slices at vdispls / rdispls, float cast when use_double is false).  Used by the CPU (gloo)
    test of the plan; the solve path never calls it."""
    import torch
    import torch.distributed as dist

    dt = np.float64 if op.use_double else np.float32
    send = v[op.vIndex].astype(dt)
    ghost = np.zeros(op.col_remote_size, dt)
    reqs = []
    recv_bufs = []
    for p, c in zip(op.recvProcRank, op.recvProcCount):
        t = torch.zeros(int(c), dtype=torch.float64 if op.use_double else torch.float32)
        recv_bufs.append((int(p), t))
        reqs.append(dist.irecv(t, src=int(p)))
    for p, c in zip(op.sendProcRank, op.sendProcCount):
        o = int(op.vdispls[p])
        reqs.append(dist.isend(torch.from_numpy(send[o:o + int(c)].copy()), dst=int(p)))
    for r in reqs:
        r.wait()
    for p, t in recv_bufs:
        o = int(op.rdispls[p])
        ghost[o:o + t.numel()] = t.numpy()
    return ghost.astype(np.float64)


def all_ranks_ok(ok: bool) -> bool:
    """True when `ok` holds on EVERY rank (one all-reduce): the way a per-rank failure becomes a decision all ranks
    take together, instead of one rank leaving a collective sequence the others are still in."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([0 if ok else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item()) == 0


def setup_p2p_halo(ctx, log=None) -> bool:
    """Switch a multi-rank context's halo exchange to NVLink peer memory: all-gather every rank's
    export blob and import them (collective).  Returns False -- on every rank, with the NCCL path left in
    place on every rank -- when the context has a single rank or when the export / import failed anywhere."""
    import torch.distributed as dist

    from .native import NativeError

    if ctx.nranks == 1:
        return False
    err = None
    try:
        mine = ctx.p2p_export()
    except NativeError as e:
        mine, err = b"", e
    blobs = [None] * ctx.nranks
    dist.all_gather_object(blobs, mine)
    ok = err is None and len(mine) > 0 and all(len(b) == len(mine) for b in blobs)
    if ok:
        try:
            ctx.p2p_import(blobs)
        except NativeError as e:
            ok, err = False, e
    if err is not None and log is not None:
        log(f"[rank {ctx.rank}] peer-memory halo unavailable: {err}")
    if not all_ranks_ok(ok):     # also the barrier after the import: nobody applies an operator before all are wired
        try:
            ctx.p2p_enable(0)
        except NativeError:
            pass
        return False
    return True
