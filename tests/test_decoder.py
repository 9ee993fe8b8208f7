"""ATRAC1 decoder on the device against the reference's TAtrac1Decoder (oracle/_ref):
bit-exact PCM for streams produced by the reference ENCODER and by this repo's encoder (long and short windows, mono
and stereo, forced window masks, low bit budgets), batch-split invariance, malformed units; plus a round-trip SNR."""
import ctypes

import numpy as np
import pytest

import atde_testlib as tl
import atracdenc_b200 as ab


def ref_decode(units, C):
    """units [F][C][212] -> pcm [F*512][C] by the reference decoder."""
    lib = tl.require_ref("ATRAC1 decoder parity")
    units = np.ascontiguousarray(units, np.uint8)
    F = units.size // (C * 212)
    pcm = np.zeros((F * 512, C), np.float32)
    lib.ref_at1_decode.restype = ctypes.c_long
    k = lib.ref_at1_decode(C, units.ctypes.data_as(tl.P), ctypes.c_long(F), pcm.ctypes.data_as(tl.P))
    assert k == F
    return pcm


def encode_streams(lib, S, F, C, seed, **kw):
    pcm = np.stack([tl.synth_rich(F, 512, C, seed=seed + s, kind=("mix", "tones", "steps")[s % 3]) if s % 2
                    else tl.synth_streams(1, F, 512, C, seed=seed + s)[0] for s in range(S)])
    enc = ab.Encoder(ab.CODEC_ATRAC1, C, lib=lib, **kw)
    units = enc.encode(pcm, S)                                   # [S][F][C][212]
    enc.close()
    return pcm, units


def check_decoder(lib, S, F, C, seed=1, cut=None, **kw):
    pcm, units = encode_streams(lib, S, F, C, seed, **kw)
    dec = ab.Decoder(C, lib=lib)
    if cut:
        got = np.concatenate([dec.decode(units[:, :cut], S), dec.decode(units[:, cut:], S)], axis=1)
    else:
        got = dec.decode(units, S)
    dec.close()
    for s in range(S):
        want = ref_decode(units[s], C)
        assert np.array_equal(got[s].view(np.uint32), want.view(np.uint32)), \
            f"stream {s}: first differing sample {np.argwhere(got[s].view(np.uint32) != want.view(np.uint32))[:3].tolist()}"
    return pcm, got


def check_malformed(lib, C=2, F=6):
    rng = np.random.default_rng(5)
    units = rng.integers(0, 256, (3, F, C, 212), dtype=np.uint8)          # random bytes: most units overrun or are invalid
    units[:, :, :, 0] &= 0x3F                                             # keep the low/mid block codes <= ... valid-ish
    units[0, :, :, 0] = 0xFC                                              # codes 3,3,3: LogCount -1, -1, 0 -> invalid -> silence
    units[1, :, :, 0] = 0xA8                                              # codes 2,2,2,0: long, long, hi LogCount 1 -> refused
    dec = ab.Decoder(C, lib=lib)
    got = dec.decode(units[0:1], 1)
    assert np.array_equal(got[0].view(np.uint32), ref_decode(units[0], C).view(np.uint32))
    dec.reset()
    with pytest.raises(ab.AtdeError):
        dec.decode(units[1:2], 1)
    dec.reset()
    units[2, :, :, 0] = np.where(rng.integers(0, 2, (F, C)) == 0, 0xA0 | 0x0C, 0x00 | 0x00)   # all-long / all-short, rest random
    got = dec.decode(units[2:3], 1)
    assert np.array_equal(got[0].view(np.uint32), ref_decode(units[2], C).view(np.uint32))
    dec.close()


def test_decoder_emulated(emu_lib):
    check_decoder(emu_lib, S=2, F=7, C=2, seed=10)
    check_decoder(emu_lib, S=1, F=6, C=1, seed=20, cut=2)
    check_decoder(emu_lib, S=1, F=5, C=2, seed=30, window_mode=0, window_mask=5)
    check_malformed(emu_lib, C=2, F=5)


@pytest.mark.gpu
def test_decoder_gpu(gpu_lib):
    pcm, got = check_decoder(gpu_lib, S=24, F=120, C=2, seed=100)
    check_decoder(gpu_lib, S=8, F=61, C=1, seed=200, cut=17)
    for mask in (1, 2, 4, 7):
        check_decoder(gpu_lib, S=4, F=33, C=2, seed=300 + mask, window_mode=0, window_mask=mask)
    check_decoder(gpu_lib, S=4, F=40, C=2, seed=400, bfu_idx_const=1)
    check_malformed(gpu_lib, C=2, F=40)
    # round trip: the decoded signal is the input delayed by the codec's analysis + synthesis latency
    x, y = pcm[1, :, 0].astype(np.float64), got[1, :, 0].astype(np.float64)
    best = max(range(0, 400), key=lambda d: float(np.dot(x[:20000], y[d:20000 + d])))
    err = y[best:best + 40000] - x[:40000]
    snr = 10 * np.log10(np.sum(x[:40000] ** 2) / np.sum(err ** 2))
    assert snr > 15.0, (best, snr)
