"""Multi-rank golden fixtures: the UNMODIFIED reference run on N MPI ranks (oracle/_ref/
libsaena_ref_mp.so over the multi-process MPI stand-in, started by oracle/mprun.py).  Frozen per rank:
its share of the hierarchy exactly as the reference laid it out (local / remote split, halo plans,
Grid::repart_u plans, shrunk coarse levels; rank numbers translated to world ranks), the reference's
own matvec / smoother / transfer outputs on slices of seeded global vectors, and its solve_pCG result.

    make -C oracle ref_mp && python tests/golden/make_golden_multirank.py
"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mprun  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, ranks, mx, what="poisson"):
    out = tempfile.mkdtemp(prefix="saena_golden_mp_")
    try:
        rc = mprun.run(ranks, [sys.executable, "-m", "oracle.mp_worker", what, str(mx), out], timeout=600,
                       env=dict(os.environ, SAENA_MP_DUMP="1", PYTHONPATH=ROOT))
        assert rc == 0, rc
        merged = {"ranks": np.array([ranks]), "mx": np.array([mx])}
        for r in range(ranks):
            d = np.load(os.path.join(out, f"rank{r}.npz"))
            for k in d.files:
                merged[f"r{r}.{k}"] = d[k]
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **merged)
        d0 = np.load(os.path.join(out, "rank0.npz"))
        print(name, "ranks", ranks, "iters", int(d0["iters"][0]), "bytes", os.path.getsize(path))
    finally:
        shutil.rmtree(out, ignore_errors=True)


if __name__ == "__main__":
    make("poisson10_np2", 2, 12)
    make("poisson14_np4", 4, 16)
    # BASELINE.json configs[4]'s synthetic unstructured shape (irregular rows, several neighbours per rank)
    make("unstructured40_np3", 3, 40, "unstructured")
