"""One MPI rank (oracle/mprun.py) of the multi-rank drop-in check: experiments/Poisson.cpp's sequence through the
public saena.hpp API on MPI_COMM_WORLD, saena::amg::solve_pCG resolved to the adaptor (this rank's GPU =
SBMPI_RANK), against saena_object::solve_pCG -- the reference's own multi-rank CPU solve -- on the same
hierarchy object.  Prints one line per rank; exits non-zero on a mismatch."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libsaena_dropin_mp.so")


def main():
    mx = int(sys.argv[1])
    L = ctypes.CDLL(LIB)
    cap = 64
    hg, hc = np.zeros(cap), np.zeros(cap)
    ig, ic, ng, nc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    du, dmv = ctypes.c_double(), ctypes.c_double()
    dp = ctypes.POINTER(ctypes.c_double)
    rc = L.dropin_poisson_check(mx, ctypes.byref(ig), ctypes.byref(ic), hg.ctypes.data_as(dp), ctypes.byref(ng),
                                hc.ctypes.data_as(dp), ctypes.byref(nc), cap, ctypes.byref(du), ctypes.byref(dmv))
    assert rc == 0
    n = min(ng.value, nc.value)
    herr = float(np.max(np.abs(hg[:n] - hc[:n]) / hc[:n]))
    print(f"rank {os.environ.get('SBMPI_RANK')}: iters gpu {ig.value} cpu {ic.value}, history err {herr:.2e}, "
          f"u err {du.value:.2e}, matvec err {dmv.value:.2e}", flush=True)
    assert abs(ig.value - ic.value) <= 1 and n >= 3
    assert herr <= 1e-6            # float_level 0: ghost values travel as float
    assert du.value < 1e-6 and dmv.value < 1e-12
    L.MPI_Finalize()


if __name__ == "__main__":
    main()
