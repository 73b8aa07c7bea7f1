"""The size-independent property checker (tests/full_size_properties.py) run through the C oracle at a size it
finishes in seconds; tests/test_zy_full_size_gpu.py runs the same checker through the CUDA path at 256^3."""
import pytest

from oracle.oracle import Oracle
from saena_b200 import sa_setup
from tests.full_size_properties import check


@pytest.mark.parametrize("n", [10, 14])
def test_properties_hold_for_the_oracle(n):
    h = sa_setup.build_hierarchy(*sa_setup.poisson3d_coo(n), device="cpu")
    out = check(Oracle(h), h, n, sa_setup.poisson3d_rhs(n))
    assert out["iterations"] <= 8
