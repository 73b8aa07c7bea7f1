"""Small driver for `ncu --set full`: builds the bench hierarchy (256^3 by default), uploads it, lets the library pick
the row mappings by measurement as bench.py does, then -- between cudaProfilerStart/Stop, so that
`ncu --profile-from-start off` sees nothing else -- runs ONE one-iteration PCG solve: 2 V-cycles (the first eager or
captured, every kernel of every level), the level-0 SpMV, the dots and the Krylov updates.  See tools/r02_call5.sh."""
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from saena_b200 import native  # noqa: E402
from saena_b200.sa_setup import build_device_hierarchy, poisson3d_coo, poisson3d_rhs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dh = build_device_hierarchy(*poisson3d_coo(n))
h = dh.to_rank(0, 1)
del dh
torch.cuda.empty_cache()
ctx = native.Context()
ctx.upload_hierarchy(h)
if "--no-autotune" not in sys.argv:
    ctx.autotune_mapping_native(10, 0.03)
ctx.set_graphs(False)   # eager launches: every kernel is its own ncu result, in V-cycle order
rhs = torch.from_numpy(poisson3d_rhs(n)).cuda()
u = torch.zeros_like(rhs)
it, hist = ctx.solve_pcg_dev(rhs.data_ptr(), u.data_ptr(), max_iter=1, tol=1e-8)   # warm
torch.cuda.synchronize()
torch.cuda.profiler.start()
it, hist = ctx.solve_pcg_dev(rhs.data_ptr(), u.data_ptr(), max_iter=1, tol=1e-8)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", it, hist, [(l, ctx.get_mapping(l, 0), ctx.get_mapping(l, 1), ctx.get_mapping(l, 2)) for l in range(len(h.levels))])
ctx.close()
