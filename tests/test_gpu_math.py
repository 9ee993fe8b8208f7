"""Device replicas of glibc's log10f/log2f/logf vs the live libm of the box (bit-exact)."""
import ctypes

import numpy as np
import pytest

import atracdenc_b200 as ab

pytestmark = pytest.mark.gpu


def _live(name, x):
    libm = ctypes.CDLL("libm.so.6")
    f = getattr(libm, name)
    f.restype = ctypes.c_float
    f.argtypes = [ctypes.c_float]
    return np.array([f(float(v)) for v in x], np.float32)


@pytest.mark.parametrize("fn,name", [(0, "log10f"), (1, "log2f"), (2, "logf")])
def test_device_math_matches_libm(gpu_lib, fn, name):
    u = np.arange(0, 0x7f800000, 8191, dtype=np.uint64).astype(np.uint32)      # 260k patterns incl. subnormals
    x = np.concatenate([u.view(np.float32), np.array([0.0, 1.0, np.inf, 1e-45], np.float32)])
    y = np.empty_like(x)
    rc = gpu_lib.atde_debug_math(0, fn, x.ctypes.data, y.ctypes.data, x.size)
    assert rc == 0, gpu_lib.atde_last_error()
    want = _live(name, x)
    assert np.array_equal(want.view(np.uint32), y.view(np.uint32))


def test_device_trig_matches_libm(gpu_lib):
    """glibc_trig.cuh on the device: sin / cos / atan (double) and sincosf against the live libm."""
    import parity_cases as pc
    pc.check_trig_replicas(gpu_lib)
