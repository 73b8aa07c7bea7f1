"""The compute side of the distributed kernels on ONE GPU (ncu cannot follow a multi-rank command):
rank `--rank`'s share of a `--ranks`-way partition of the bench hierarchy in a detached context
(saena_b200_init_detached: no peers; ghost values are zeros), every operator that has a halo timed
compute-only (no pack, no flags) through the fused halo kernel and through the separate interior +
boundary kernels.

    python tools/profile_fused.py --n 256 --ranks 4 --rank 1
    ncu --set full --clock-control none --import-source on -k regex:fused_halo_spmv -c 8 \\
        -o gpurun_out/r02_fused python tools/profile_fused.py --n 256 --ranks 4 --rank 1 --only 1:A
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from saena_b200.hierarchy import KIND_A, KIND_P, KIND_R  # noqa: E402
from saena_b200.native import Context  # noqa: E402
from saena_b200.sa_setup import build_device_hierarchy, poisson3d_coo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--ranks", type=int, default=4)
    ap.add_argument("--rank", type=int, default=1)
    ap.add_argument("--agglomerate-below", type=int, default=10_000)
    ap.add_argument("--rebalance-above", type=float, default=1.10)
    ap.add_argument("--only", default="", help="level:kind pairs, e.g. 1:A,1:P")
    a = ap.parse_args()
    only = {(int(x.split(":")[0]), "APR".index(x.split(":")[1])) for x in a.only.split(",") if x}
    dh = build_device_hierarchy(*poisson3d_coo(a.n))
    h = dh.to_rank(a.rank, a.ranks, agglomerate_below=a.agglomerate_below, rebalance_above=a.rebalance_above)
    del dh
    ctx = Context(rank=a.rank, nranks=a.ranks, detached=True)
    ctx.upload_hierarchy(h)
    for l, lv in enumerate(h.levels):
        for k, op in ((KIND_A, lv.A), (KIND_P, lv.P), (KIND_R, lv.R)):
            if op is None or op.M == 0 or op.nnz_remote == 0 or (only and (l, k) not in only):
                continue
            nbytes = ctx.operator_bytes(l, k)
            big = nbytes > 300e6
            f = ctx.time_matvec_compute_only(l, k, True, 20, flush_l2=not big)
            s = ctx.time_matvec_compute_only(l, k, False, 20, flush_l2=not big)
            print(json.dumps({"level": l, "kind": "APR"[k], "rows": op.M, "nnz": op.nnz, "ghost_values": op.recvSize,
                              "mapping": ctx.get_mapping(l, k), "fused_ms": round(f, 4), "separate_ms": round(s, 4),
                              "fused_GBs": round(nbytes / f / 1e6, 1), "separate_GBs": round(nbytes / s / 1e6, 1)}),
                  flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
