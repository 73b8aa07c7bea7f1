"""BASELINE.json configs[3]: standalone saena_matrix::matvec and smoother-sweep bandwidth sweep on the
band pattern of experiments/banded.cpp (row i = columns [i-b, i+b], value 1/(i+j+1)), generated on the
device (saena_b200_upload_band_operator) up to the full 50 M rows x 129 entries = 77 GB, 64-bit row
offsets.  Per point: CUDA-event time of 20 launches of the SpMV and of the fused Chebyshev sweep,
algorithmic GB/s (12 nnz + 20 M, offsets 8 B when nnz >= 2^31; first sweep + 24 M), and a parity
property that holds at any size: sampled rows of A*1 and of A*v (v seeded uniform(-1,1)) against the
band sums written out on the host.

    python tools/band_sweep.py [--sizes 1000000,10000000,50000000] [--bands 0,1,2,4,8,16,32,64]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from saena_b200.hierarchy import KIND_A  # noqa: E402
from saena_b200.native import Context  # noqa: E402

HBM_BYTES = 178e9


def band_rows(n, b, v, rows):
    out = np.empty(len(rows))
    for k, i in enumerate(rows):
        j = np.arange(max(i - b, 0), min(i + b, n - 1) + 1)
        out[k] = np.sum(v[j] / (i + j + 1.0))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1000000,10000000,50000000")
    ap.add_argument("--bands", default="0,1,2,4,8,16,32,64")
    ap.add_argument("--peak", type=float, default=None)
    a = ap.parse_args()
    peak = a.peak
    if peak is None:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        peak = float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0
    ctx = Context()
    rng = np.random.default_rng(12345)
    for n in (int(x) for x in a.sizes.split(",")):
        v = rng.uniform(-1, 1, n)
        ones = np.ones(n)
        rows = np.unique(np.concatenate((np.arange(min(n, 70)), np.arange(max(n - 70, 0), n),
                                         rng.integers(0, n, 400))))
        for b in (int(x) for x in a.bands.split(",")):
            if b >= n:
                continue
            nnz_est = n * (2 * b + 1)
            csr = 12 * nnz_est + 8 * n
            vectors = 14 * 8 * n
            # the sliced copy doubles the operator: beyond half the HBM the CSR entries are released
            # once it is built (CSR + sliced copy of the full-size case peak at 156 GB + vectors)
            big = 2 * csr + vectors > 0.6 * HBM_BYTES
            t = time.time()
            nnz = ctx.upload_band(n, b, eig_max=2.0, sliced_only=big)
            setup_s = time.time() - t
            got1, gotv = ctx.matvec(0, KIND_A, ones), ctx.matvec(0, KIND_A, v)
            e1 = float(np.max(np.abs(got1[rows] - band_rows(n, b, ones, rows)) / np.abs(band_rows(n, b, ones, rows))))
            wv = band_rows(n, b, v, rows)
            ev = float(np.linalg.norm(gotv[rows] - wv) / np.linalg.norm(wv))
            assert e1 <= 1e-12 and ev <= 1e-12, (n, b, e1, ev)
            p = 8 if nnz >= 2 ** 31 - 1 else 4
            mv_bytes = 12 * nnz + (p + 16) * n
            sw_bytes = mv_bytes + 24 * n          # first sweep: + rhs, inv_diag in, d out (u out is the SpMV's w)
            ctx.time_matvec(0, KIND_A, 3)
            mv = ctx.time_matvec(0, KIND_A, 20, flush_l2=mv_bytes < 300e6)
            sw = ctx.time_smooth_sweep(0, "chebyshev", 20, flush_l2=mv_bytes < 300e6)
            print(json.dumps({"n": n, "half_bandwidth": b, "nnz": nnz, "nnz_per_row": round(nnz / n, 2),
                              "row_offsets": "int64" if p == 8 else "int32", "mapping": ctx.get_mapping(0, KIND_A),
                              "spmv_ms": round(mv, 4), "spmv_GBs": round(mv_bytes / mv / 1e6, 1),
                              "spmv_frac": round(mv_bytes / mv / 1e6 / peak, 3),
                              "cheb_sweep_ms": round(sw, 4), "cheb_sweep_GBs": round(sw_bytes / sw / 1e6, 1),
                              "cheb_sweep_frac": round(sw_bytes / sw / 1e6 / peak, 3),
                              "err_rows_A1": e1, "err_rows_Av": ev, "setup_s": round(setup_s, 2)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
