"""Restriction fused with the residual (north_star bullet) -- the measurement: on the bench hierarchy (256^3 by
default), levels whose A runs on the sliced layout, the solve path's two kernels (residual written to res, then R)
against the one-kernel scatter form (csrc/fused_restrict.cu).  One JSON line per level.

    python tools/fused_restrict_bench.py [n] > gpurun_out/r02_fused_restrict.jsonl
"""
import json
import sys

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import torch  # noqa: E402

from saena_b200 import native  # noqa: E402
from saena_b200.hierarchy import KIND_A, KIND_P, KIND_R  # noqa: E402
from saena_b200.sa_setup import build_device_hierarchy, poisson3d_coo  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dh = build_device_hierarchy(*poisson3d_coo(n))
h = dh.to_rank(0, 1)
del dh
torch.cuda.empty_cache()
ctx = native.Context()
ctx.upload_hierarchy(h)
rng = np.random.default_rng(3)
for l, lv in enumerate(h.levels[:-1]):
    if ctx.get_mapping(l, KIND_A) != 100:
        continue
    u, b = rng.uniform(-1, 1, lv.A.M), rng.uniform(-1, 1, lv.A.M)
    two, fused, diff = ctx.time_residual_restrict(l, u, b, 20)
    a_bytes, r_bytes = ctx.operator_bytes(l, KIND_A), ctx.operator_bytes(l, KIND_R)
    print(json.dumps({"level": l, "fine_rows": lv.A.M, "coarse_rows": lv.R.M, "nnz_A": lv.A.nnz, "nnz_P": lv.P.nnz,
                      "two_kernels_ms": round(two, 4), "fused_scatter_ms": round(fused, 4),
                      "fused_over_two": round(fused / two, 3), "rel_diff": diff,
                      "bytes_saved_by_fusing": 16 * lv.A.M, "atomics_added": int(lv.P.nnz),
                      "two_kernels_algorithmic_GBs": round((a_bytes + 8 * lv.A.M + r_bytes) / two / 1e6, 1)}), flush=True)
ctx.close()
