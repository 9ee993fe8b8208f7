"""Parity checks shared by the CPU-emulation suite (tests/test_emu_parity.py, tiny sizes) and the
GPU suite (tests/test_gpu_parity.py).  Every function takes a loaded C-ABI library (ctypes) and
calls the product through include/atde_b200.h only.  The ORACLE side is tests/atde_testlib.py."""
from pathlib import Path

import numpy as np

import atde_testlib as tl
import atracdenc_b200 as ab

GOLDEN = Path(__file__).resolve().parent / "golden"


def oracle_at1(C, pcm_stream, **kw):
    """Reference output for one stream, [F][C][212] + sizes: oracle/_ref when it travelled with the
    repo, else the C restatement (itself pinned against _ref by test_oracle_vs_ref.py)."""
    F = pcm_stream.size // C // 512
    if tl.ref_lib() is not None:
        payload, sizes = tl.ref_encode(1, C, pcm_stream, **kw)
        return tl.pad_units(payload, sizes, 212)[:F * C].reshape(F, C, 212), sizes[:F * C].reshape(F, C)
    units, sizes = tl.port_at1_encode(C, pcm_stream, **kw)
    return units.reshape(F, C, 212), sizes.reshape(F, C)


def check_at1_golden_config1(lib):
    g = np.load(GOLDEN / "at1_config1_sine_mono.npz")
    frames = tl.engine_view(tl.config1_sine(), 1, 512, total=44100)
    enc = ab.Encoder(ab.CODEC_ATRAC1, 1, lib=lib)
    out, sizes = enc.encode(frames, 1, want_sizes=True)
    enc.close()
    assert out.shape == (1, 88, 1, 212)
    assert np.array_equal(out.reshape(88, 212), g["units"])
    assert np.array_equal(sizes.reshape(-1), g["sizes"])


def check_at1_golden_stereo(lib, max_frames=None):
    g = np.load(GOLDEN / "at1_stereo_bursts.npz")
    pcm, units, sizes, masks = g["pcm"], g["units"], g["sizes"], g["masks"]
    S, F = units.shape[0], units.shape[1]
    if max_frames:
        F = min(F, max_frames)
        pcm = pcm[:, :F * 512]
    enc = ab.Encoder(ab.CODEC_ATRAC1, 2, lib=lib)
    enc.arm_taps()
    out, got_sizes = enc.encode(pcm, S, want_sizes=True)
    got_masks = enc.tap(ab.TAP_MASKS, (S, F, 2), np.uint8)
    enc.close()
    assert np.array_equal(got_masks, masks[:, :F])
    assert np.array_equal(got_sizes, sizes[:, :F])
    assert np.array_equal(out, units[:, :F])


def check_at1_vs_oracle(lib, S, F, C, seed=1, **settings):
    pcm = tl.synth_streams(S, F, 512, C, seed=seed)
    kw = {}
    enc_kw = {}
    if "window_mask" in settings:
        kw.update(window_auto=0, window_mask=settings["window_mask"])
        enc_kw.update(window_mode=0, window_mask=settings["window_mask"])
    if "bfu" in settings:
        kw.update(bfu=settings["bfu"])
        enc_kw.update(bfu_idx_const=settings["bfu"])
    enc = ab.Encoder(ab.CODEC_ATRAC1, C, lib=lib, **enc_kw)
    out, sizes = enc.encode(pcm, S, want_sizes=True)
    enc.close()
    for s in range(S):
        want, want_sizes = oracle_at1(C, pcm[s].reshape(-1), **kw)
        assert np.array_equal(sizes[s], want_sizes), f"stream {s}: payload sizes differ"
        bad = np.argwhere((out[s] != want).any(-1))
        assert bad.size == 0, f"stream {s}: first differing (frame, channel) = {bad[:4].tolist()}"


def check_at1_stage_taps(lib, S=2, F=12, C=2):
    """Intermediates against the reference's own sub-objects (needs oracle/_ref)."""
    if tl.ref_lib() is None:
        return False
    pcm = tl.synth_streams(S, F, 512, C, seed=3)
    enc = ab.Encoder(ab.CODEC_ATRAC1, C, lib=lib)
    enc.arm_taps()
    enc.encode(pcm, S)
    specs = enc.tap(ab.TAP_SPECS, (S, F, C, 512), np.float32)
    masks = enc.tap(ab.TAP_MASKS, (S, F, C), np.uint8)
    chl = enc.tap(ab.TAP_CHLOUD, (S, F, C), np.float32)
    loud = enc.tap(ab.TAP_LOUDNESS, (S, F), np.float32)
    sfi = enc.tap(ab.TAP_SFI, (S, F, C, 52), np.uint8)
    enc.close()
    for s in range(S):
        st = tl.ref_at1_stages(C, pcm[s].reshape(-1))
        assert np.array_equal(st["masks"], masks[s])
        assert np.array_equal(st["specs"].view(np.uint32), specs[s].view(np.uint32))
        assert np.array_equal(st["chloud"].view(np.uint32), chl[s].view(np.uint32))
        assert np.array_equal(st["loud"].view(np.uint32), loud[s].view(np.uint32))
        assert np.array_equal(st["sfi"], sfi[s])
    return True


def check_at1_batch_split_invariance(lib, S=3, F=16, C=2, cut=5):
    """Streams continue across calls: encoding [0,cut) then [cut,F) must equal one batch of F
    (exercises the PCM-halo / loudness carry)."""
    pcm = tl.synth_streams(S, F, 512, C, seed=5)
    enc = ab.Encoder(ab.CODEC_ATRAC1, C, lib=lib)
    whole = enc.encode(pcm, S)
    enc.reset()
    a = enc.encode(pcm[:, :cut * 512], S)
    b = enc.encode(pcm[:, cut * 512:], S)
    enc.close()
    assert np.array_equal(np.concatenate([a, b], axis=1), whole)


def check_at1_stream_independence(lib, C=2, F=8):
    """A stream's bitstream does not depend on its neighbours in the batch or its position."""
    pcm = tl.synth_streams(4, F, 512, C, seed=9)
    enc = ab.Encoder(ab.CODEC_ATRAC1, C, lib=lib)
    batch = enc.encode(pcm, 4)
    enc.reset()
    rev = enc.encode(pcm[::-1].copy(), 4)
    enc.reset()
    one = enc.encode(pcm[2:3].copy(), 1)
    enc.close()
    assert np.array_equal(batch, rev[::-1])
    assert np.array_equal(batch[2:3], one)


def check_at1_edge_inputs(lib):
    """Silence, digital full scale, alternating +-1 (Nyquist), single impulse, mono; ragged tail
    (F not a multiple of the kernel's frame tile)."""
    F = 7
    n = F * 512
    cases = {
        "silence": np.zeros((n, 2), np.float32),
        "full_scale_dc": np.full((n, 2), 32767 / 32768, np.float32),
        "nyquist": np.tile(np.array([[1.0], [-1.0]], np.float32), (n // 2, 2)) * np.float32(32767 / 32768),
        "impulse": np.zeros((n, 2), np.float32),
    }
    cases["impulse"][700, 0] = -1.0
    for name, x in cases.items():
        enc = ab.Encoder(ab.CODEC_ATRAC1, 2, lib=lib)
        out, sizes = enc.encode(x, 1, want_sizes=True)
        enc.close()
        want, want_sizes = oracle_at1(2, x.reshape(-1))
        assert np.array_equal(out[0], want), name
        assert np.array_equal(sizes[0], want_sizes), name
    mono = tl.synth_streams(1, 5, 512, 1, seed=11)
    enc = ab.Encoder(ab.CODEC_ATRAC1, 1, lib=lib)
    out = enc.encode(mono, 1)
    enc.close()
    want, _ = oracle_at1(1, mono.reshape(-1))
    assert np.array_equal(out[0], want)


def check_errors(lib):
    import pytest
    enc = ab.Encoder(ab.CODEC_ATRAC1, 2, lib=lib)
    with pytest.raises(ab.AtdeError):
        enc._check(lib.atde_encode_batch(enc.h, None, 1, 1, None, None))
    x = np.zeros(512 * 2, np.float32)
    out = np.zeros(2 * 212, np.uint8)
    with pytest.raises(ab.AtdeError):
        enc._check(lib.atde_encode_batch(enc.h, x.ctypes.data, 0, 1, out.ctypes.data, None))     # empty batch
    enc.encode(np.zeros((2, 512, 2), np.float32), 2)
    with pytest.raises(ab.AtdeError):                                                           # stream count changed without reset
        enc.encode(np.zeros((3, 512, 2), np.float32), 3)
    enc.close()
    with pytest.raises(ab.AtdeError):
        ab.Encoder(ab.CODEC_ATRAC1, 3, lib=lib)
    with pytest.raises(ab.AtdeError):
        ab.Encoder(99, 2, lib=lib)
