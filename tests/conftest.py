import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libsaena_ref.so (the compiled reference)")


def _cuda_ok() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a GPU test on a box without a GPU is an error of the invocation, not a skip: the driver
    # selects with -m "not gpu" / -m gpu.  Only the compiled-reference marker degrades to skip.
    from oracle import ref
    if not ref.available():
        skip = pytest.mark.skip(reason="oracle/_ref/libsaena_ref.so not built (needs /root/reference)")
        for item in items:
            if "ref" in item.keywords:
                item.add_marker(skip)
