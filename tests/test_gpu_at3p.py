"""GPU parity suite, ATRAC3plus stage kernels (PQF analysis, MDCT-256 x16) against the
reference's taps (oracle/_ref)."""
import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu


def test_pqf(gpu_lib):
    assert pc.check_at3p_pqf(gpu_lib, S=8, F=24, C=2) in (0, 8)
    pc.check_at3p_pqf(gpu_lib, S=3, F=10, C=1, seed=905)


def test_mdct(gpu_lib):
    pc.check_at3p_mdct(gpu_lib, S=6, F=20, C=2)
    pc.check_at3p_mdct(gpu_lib, S=2, F=9, C=1, seed=915)
