"""Pins the plain-C oracle restatement (oracle/*.c) bit-for-bit against the reference itself
(oracle/_ref = unmodified reference sources compiled here).  The reference's own tests hold no
bit-exact vectors for this path, so the reference's output is the pin."""
import numpy as np
import pytest

import atde_testlib as tl


def signals():
    rng = np.random.default_rng(1)
    n = np.arange(4096 * 11)
    yield "sine_mono", 1, tl.quantise(0.5 * np.sin(2 * np.pi * 1000 * n / 44100))
    F = 64
    x = 0.25 * rng.uniform(-1, 1, (F * 512, 2)) + 0.3 * np.sin(2 * np.pi * 440 * np.arange(F * 512) / 44100)[:, None]
    for f in range(0, F, 7):
        x[f * 512:f * 512 + 64] *= 8
    yield "noise_burst_stereo", 2, tl.quantise(x).ravel()
    yield "silence", 2, np.zeros(8192 * 2, np.float32)
    yield "quiet", 2, tl.quantise(1e-3 * rng.uniform(-1, 1, (4096 * 4, 2))).ravel()
    yield "full_scale", 1, tl.quantise(rng.uniform(-1, 1, 4096 * 4))


@pytest.mark.parametrize("kw", [{}, {"window_auto": 0, "window_mask": 0}, {"window_auto": 0, "window_mask": 7},
                                {"window_auto": 0, "window_mask": 5}, {"bfu": 3}, {"bfu": 8}])
def test_at1_port_matches_reference(ref, kw):
    for name, ch, pcm in signals():
        payload, sizes = tl.ref_encode(1, ch, pcm, **kw)
        want = tl.pad_units(payload, sizes, 212)
        got, got_sizes = tl.port_at1_encode(ch, pcm, **kw)
        assert want.shape == got.shape, name
        assert np.array_equal(sizes, got_sizes), name
        assert np.array_equal(want, got), name


def test_tables_match_reference(ref):
    P = tl.P
    a = [np.zeros(n, np.float32) for n in (48, 32, 64, 512, 52)]
    b = [np.zeros(n, np.float32) for n in (48, 32, 64, 512, 52)]
    ref.ref_tables_at1(*[x.ctypes.data_as(P) for x in a])
    tl.port_lib().oat1_tables(*[x.ctypes.data_as(P) for x in b])
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


@pytest.mark.parametrize("n", [8, 16, 64, 128, 256, 2048])
def test_kissfft_restatement(ref, n):
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, 2 * n).astype(np.float32)
    port = tl.port_lib()
    for inverse in (0, 1):
        want = np.zeros(2 * n, np.float32)
        ref.ref_kiss_fft(n, inverse, x.ctypes.data_as(tl.P), want.ctypes.data_as(tl.P))
        port.okiss_alloc.restype = tl.P
        st = port.okiss_alloc(n, inverse)
        got = np.zeros(2 * n, np.float32)
        port.okiss_fft(tl.P(st), x.ctypes.data_as(tl.P), got.ctypes.data_as(tl.P))
        port.okiss_free(tl.P(st))
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


@pytest.mark.parametrize("n", [512, 4096])
def test_kiss_real_fft_restatement(ref, n):
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    port = tl.port_lib()
    port.okissr_alloc.restype = tl.P
    want = np.zeros(n + 2, np.float32)
    got = np.zeros(n + 2, np.float32)
    ref.ref_kiss_fftr(n, x.ctypes.data_as(tl.P), want.ctypes.data_as(tl.P))
    st = port.okissr_alloc(n, 0)
    port.okiss_fftr(tl.P(st), x.ctypes.data_as(tl.P), got.ctypes.data_as(tl.P))
    port.okissr_free(tl.P(st))
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
    back_w = np.zeros(n, np.float32)
    back_g = np.zeros(n, np.float32)
    ref.ref_kiss_fftri(n, want.ctypes.data_as(tl.P), back_w.ctypes.data_as(tl.P))
    st = port.okissr_alloc(n, 1)
    port.okiss_fftri(tl.P(st), want.ctypes.data_as(tl.P), back_g.ctypes.data_as(tl.P))
    port.okissr_free(tl.P(st))
    assert np.array_equal(back_w.view(np.uint32), back_g.view(np.uint32))


@pytest.mark.parametrize("n,scale", [(512, 1.0), (256, 0.5), (64, 0.5)])
def test_mdct_restatement(ref, n, scale):
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    want = np.zeros(n // 2, np.float32)
    ref.ref_mdct(n, tl.ctypes.c_float(scale), x.ctypes.data_as(tl.P), want.ctypes.data_as(tl.P))
    port = tl.port_lib()
    port.omdct_alloc.restype = tl.P
    m = port.omdct_alloc(n, tl.ctypes.c_float(scale))
    got = np.zeros(n // 2, np.float32)
    port.omdct_run(tl.P(m), x.ctypes.data_as(tl.P), got.ctypes.data_as(tl.P))
    port.omdct_free(tl.P(m))
    assert np.array_equal(want.view(np.uint32), got.view(np.uint32))


def test_mdct_vs_naive_dft():
    """Restates src/lib/mdct/mdct_ut.cpp:74-228: MDCT against the O(N^2) cosine sum, eps = magnitude*10^(-114/20)."""
    port = tl.port_lib()
    port.omdct_alloc.restype = tl.P
    for n in (32, 64, 128, 256):
        rng = np.random.Generator(np.random.MT19937(0x4d443254))
        x = rng.uniform(-32768, 32767, n).astype(np.float32)
        m = port.omdct_alloc(n, tl.ctypes.c_float(n))          # scale N => plain cosine sum, as the ut does
        got = np.zeros(n // 2, np.float32)
        port.omdct_run(tl.P(m), x.ctypes.data_as(tl.P), got.ctypes.data_as(tl.P))
        port.omdct_free(tl.P(m))
        k = np.arange(n // 2)[:, None]
        j = np.arange(n)[None, :]
        naive = (x.astype(np.float64)[None, :] * np.cos(np.pi / (n / 2) * (j + 0.5 + n / 4) * (k + 0.5))).sum(1)
        eps = 32768.0 * n * 10 ** (-114 / 20)
        assert np.max(np.abs(got - naive)) < max(eps, 1e-3 * np.max(np.abs(naive)))
