"""Small driver for `ncu --set full`: builds the bench hierarchy (256^3 by default), uploads it and
runs two one-iteration PCG solves (= 2 x 2 V-cycles), nothing else.  See tools/profile_full.sh."""
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
rhs = torch.from_numpy(poisson3d_rhs(n)).cuda()
u = torch.zeros_like(rhs)
for _ in range(2):
    it, hist = ctx.solve_pcg_dev(rhs.data_ptr(), u.data_ptr(), max_iter=1, tol=1e-8)
print("ok", it, hist)
ctx.close()
