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


def cpu_baseline(n: int, b: int, reps: int = 3):
    """The CPU side of the same point (the reference's CPU path timed beside it).  The reference's
    own assembler cannot build this matrix at size (std::set inserts, and saena::band_matrix's fill is compiled out,
    src/aux_functions2.cpp:1344-1371), so the C restatement of saena_matrix::matvec (oracle/saena_oracle.c:so_matvec,
    pinned against the reference on the band pattern by golden band8_1500) runs on the arrays written out directly:
    kind "port", one core -- the loop of src/saena_matrix_matvec.cpp:55-80 is single-threaded in the reference too.
    The fused smoother sweep has no CPU counterpart: the reference's chebyshev is the SpMV + two vector passes."""
    from oracle.oracle import Oracle
    from saena_b200.hierarchy import Hierarchy, Level, Operator
    i = np.arange(n, dtype=np.int64)
    lo, hi = np.maximum(i - b, 0), np.minimum(i + b, n - 1)
    counts = (hi - lo + 1).astype(np.int32)
    rows = np.repeat(i, counts)
    first = np.cumsum(counts) - counts
    cols = (np.arange(len(rows), dtype=np.int64) - np.repeat(first, counts)) + np.repeat(lo, counts)
    vals = 1.0 / (rows + cols + 1.0)
    op = Operator(kind=KIND_A, level=0, M=n, Mbig=n, Nbig=n, row_offset=0, col_offset=0, n_local_cols=n,
                  nnzPerRow_local=counts, col_local=cols.astype(np.int32), val_local=vals)
    h = Hierarchy([Level(0, op, inv_diag=2.0 * i + 1.0, eig_max=2.0)], coarse_n=0)
    o = Oracle(h)
    v = np.random.default_rng(12345).uniform(-1, 1, n)
    w = o.matvec(0, KIND_A, v)
    sample = np.unique(np.concatenate((np.arange(min(n, 20)), np.arange(max(n - 20, 0), n))))
    assert np.allclose(w[sample], band_rows(n, b, v, sample), rtol=1e-12)
    t = time.perf_counter()
    for _ in range(reps):
        o.matvec(0, KIND_A, v)
    sec = (time.perf_counter() - t) / reps
    nnz = int(counts.sum())
    mv_bytes = 12 * nnz + 20 * n
    print(json.dumps({"cpu_baseline": {"value": round(mv_bytes / sec / 1e9, 2), "unit": "GB/s", "cores": 1, "kind": "port",
                                       "sample": f"oracle so_matvec (C restatement of saena_matrix::matvec) on the band "
                                                 f"pattern, {n} rows, half bandwidth {b}, {nnz} non-zeros; {reps} "
                                                 f"applications, {sec * 1e3:.1f} ms each"},
                      "n": n, "half_bandwidth": b, "nnz": nnz, "spmv_ms": round(sec * 1e3, 2)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-only", action="store_true", help="only the CPU baseline points (no GPU needed)")
    ap.add_argument("--cpu-sizes", default="1000000", help="rows of the CPU baseline points (half bandwidth 64 and 1)")
    ap.add_argument("--sizes", default="1000000,10000000,50000000")
    ap.add_argument("--bands", default="0,1,2,4,8,16,32,64")
    ap.add_argument("--peak", type=float, default=None)
    a = ap.parse_args()
    peak = a.peak
    if peak is None:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        peak = float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0
    for n in (int(x) for x in filter(None, a.cpu_sizes.split(","))):
        for b in (64, 1):
            cpu_baseline(n, b)
    if a.cpu_only:
        return
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
