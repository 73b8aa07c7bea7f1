"""The drop-in adaptor's upload walk on several MPI ranks, without a GPU.

saena_b200/adaptor/saena_b200_adaptor.cpp walks saena_object::grids on every rank, translates the rank numbers
of each level's (shrinking) communicator to world ranks, uploads empty operators on ranks a shrink left out of a
level, and hands everything to the C ABI.  Here the ABI is a recorder (oracle/abi_recorder.cpp) linked with the
reference (multi-process MPI stand-in) and the adaptor: one solve through the public API
(saena::amg::solve_pCG) on N ranks, then every recorded array is compared with what oracle/ref.py extracts from
the same solver object -- the extraction the multi-rank oracle is pinned with against the reference's own numbers
(tests/test_multirank_reference.py)."""
import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.ref


@pytest.mark.parametrize("ranks,what,size", [(1, "poisson", 12), (2, "poisson", 14), (4, "poisson", 20),
                                             (7, "poisson", 24), (3, "unstructured", 40)])
def test_adaptor_uploads_what_the_reference_laid_out(ranks, what, size, tmp_path):
    from oracle import mprun, ref
    if not os.path.exists(ref.REC_LIB_PATH):
        pytest.skip("oracle/_ref/libsaena_dropin_rec_mp.so not built (make -C oracle dropin_rec_mp)")
    out = str(tmp_path / "rec")
    rc, outs = mprun.run(ranks, [sys.executable, "-m", "oracle.mp_worker", what, str(size), out], timeout=600,
                         env=dict(os.environ, SAENA_MP_ADAPTOR_CHECK="1", SAENA_REF_LIB_PATH=ref.REC_LIB_PATH,
                                  PYTHONPATH=ROOT), capture=True)
    assert rc == 0, "\n".join(o[-1500:] for o in outs)
    assert len(glob.glob(os.path.join(out, "adaptor_ok_*"))) == ranks


@pytest.mark.parametrize("ranks", [1, 3])
def test_adaptor_uploads_again_after_a_lazy_update(ranks, tmp_path):
    """a lazy update (saena::amg::update1, src/saena_object_lazy.cpp:7-35 -- compiled out in this version of the
    reference, so its effect is applied directly) replaces grids[0].A with a matrix of other values; the drop-in
    keeps its device copy keyed by the solver object, so it must notice and upload again"""
    from oracle import mprun, ref
    if not os.path.exists(ref.REC_LIB_PATH):
        pytest.skip("oracle/_ref/libsaena_dropin_rec_mp.so not built (make -C oracle dropin_rec_mp)")
    out = str(tmp_path / "rec")
    rc, outs = mprun.run(ranks, [sys.executable, "-m", "oracle.mp_worker", "poisson", "14", out], timeout=600,
                         env=dict(os.environ, SAENA_MP_ADAPTOR_CHECK="1", SAENA_MP_ADAPTOR_UPDATE="14",
                                  SAENA_REF_LIB_PATH=ref.REC_LIB_PATH, PYTHONPATH=ROOT), capture=True)
    assert rc == 0, "\n".join(o[-1500:] for o in outs)
    assert len(glob.glob(os.path.join(out, "adaptor_ok_*"))) == ranks
