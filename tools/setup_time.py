"""Setup time of the bench hierarchy (256^3 by default) with the two product routes of saena_b200/sa_setup.py: the
library's device SpGEMM (csrc/spgemm.cu) and the tensor-op expand / sort / compress route; level sizes must agree.

    python tools/setup_time.py [n] > gpurun_out/r02_setup_time.json
"""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

from saena_b200 import sa_setup  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
out = {"n": n}
sizes = {}
for route in ("native", "torch"):
    os.environ["SAENA_SETUP_SPGEMM"] = route
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    t = time.perf_counter()
    try:
        dh = sa_setup.build_device_hierarchy(*sa_setup.poisson3d_coo(n), device="cuda")
        torch.cuda.synchronize()
        out[route] = {"setup_s": round(time.perf_counter() - t, 2), "levels": len(dh.levels),
                      "peak_torch_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1)}
        sizes[route] = [(lv.A.n_rows, lv.A.nnz) for lv in dh.levels]
        del dh
    except Exception as e:   # report, go on with the other route
        out[route] = {"error": repr(e)[:300]}
    torch.cuda.empty_cache()
out["level_rows_nnz"] = sizes.get("native") or sizes.get("torch")
out["same_level_sizes"] = sizes.get("native") == sizes.get("torch") if len(sizes) == 2 else None
print(json.dumps(out))
