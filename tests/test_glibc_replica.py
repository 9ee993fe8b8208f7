"""The per-frame libm calls of the reference (log10f, log2f) resolve to glibc 2.39; the oracle's
restatement (oracle/glibc_replica.c) must equal the LIVE libm of this box bit for bit.  The full
2^31-point sweep lives in tests/tools/glibc_exhaustive.c (30 s on 8 cores); here a strided sweep
plus the ranges the encoder actually visits."""
import ctypes

import numpy as np

import atde_testlib as tl


def _live(name):
    libm = ctypes.CDLL("libm.so.6")
    f = getattr(libm, name)
    f.restype = ctypes.c_float
    f.argtypes = [ctypes.c_float]
    return f


def _check(name, port_name, xs):
    live, port = _live(name), getattr(tl.port_lib(), port_name)
    for x in xs:
        a = np.float32(live(float(x)))
        b = np.float32(port(float(x)))
        assert a.view(np.uint32) == b.view(np.uint32) or (np.isnan(a) and np.isnan(b)), (name, x, a, b)


def _samples():
    u = np.arange(0, 0x7f800000, 104729, dtype=np.uint64).astype(np.uint32)
    xs = list(u.view(np.float32))
    rng = np.random.default_rng(7)
    xs += list(np.exp(rng.uniform(-30, 3, 3000)).astype(np.float32))
    xs += [np.float32(v) for v in (0.0, 1.0, 2.0, 0.5, 1e-38, 1e-45, 3.4e38, np.inf)]
    return xs


def test_log10f_matches_libm():
    _check("log10f", "og_log10f", _samples())


def test_log2f_matches_libm():
    _check("log2f", "og_log2f", _samples())


def test_logf_matches_libm():
    _check("logf", "og_logf", _samples())
