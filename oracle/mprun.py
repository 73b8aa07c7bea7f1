"""mprun -- starts N processes of one command as the ranks of the multi-process MPI stand-in
(oracle/ref_shim_mp/mpi_multi.c).  TEST INFRASTRUCTURE: lets the unmodified reference run its real
multi-rank code paths in an image without MPI.

    python -m oracle.mprun -n 4 python some_script.py args...

Every rank gets SBMPI_RANK / SBMPI_SIZE / SBMPI_DIR (a private temporary directory holding the
rendezvous sockets).  The exit code is the first non-zero exit code of a rank; a rank that dies
takes the others down (they see its socket close)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
import time


def run(n: int, argv: list[str], timeout: float | None = None, env: dict | None = None, capture: bool = False):
    d = tempfile.mkdtemp(prefix="sbmpi_")
    procs = []
    try:
        for r in range(n):
            e = dict(os.environ if env is None else env, SBMPI_RANK=str(r), SBMPI_SIZE=str(n), SBMPI_DIR=d)
            procs.append(subprocess.Popen(argv, env=e, stdout=subprocess.PIPE if capture else None,
                                          stderr=subprocess.STDOUT if capture else None, text=capture))
        deadline = None if timeout is None else time.time() + timeout
        rc, outs = 0, [""] * n
        pending = set(range(n))
        while pending:
            for r in list(pending):
                code = procs[r].poll()
                if code is not None:
                    pending.discard(r)
                    if capture:
                        outs[r] = procs[r].stdout.read()
                    if code and not rc:
                        rc = code
            if rc or (deadline and time.time() > deadline):
                if not rc:
                    rc = 124
                time.sleep(0.3)
                for r in pending:
                    procs[r].kill()
                for r in list(pending):
                    procs[r].wait()
                    if capture:
                        outs[r] = procs[r].stdout.read()
                pending.clear()
            time.sleep(0.01)
        return (rc, outs) if capture else rc
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    a = sys.argv[1:]
    if len(a) < 3 or a[0] != "-n":
        raise SystemExit(__doc__)
    raise SystemExit(run(int(a[1]), a[2:]))


if __name__ == "__main__":
    main()
