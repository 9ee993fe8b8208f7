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
    tl.require_ref()
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


# =================================================================================================
# ATRAC3 (LP2 / LP4).  Oracle = the reference encoder itself (oracle/_ref, driven frame by frame
# through TAtrac3Encoder::GetLambda by oracle/ref_harness_at3.cpp); where it did not travel, the
# committed fixtures under tests/golden/ (generated from it by make_golden.py) are the pin.
def oracle_at3(C, pcm_stream, kbit=0, no_gain=0, no_tonal=0):
    """Reference frames [Fo][FrameSz] for one stream (skips / fails the calling test without oracle/_ref)."""
    tl.require_ref()
    return tl.ref_at3_stages(C, pcm_stream, kbit, no_gain, no_tonal)[2]


def _at3_enc(lib, C, kbit=0, no_gain=0, no_tonal=0, **kw):
    return ab.Encoder(ab.CODEC_ATRAC3, C, bitrate=kbit * 1024, no_gain_control=bool(no_gain),
                      no_tonal=bool(no_tonal), lib=lib, **kw)


def check_at3_golden(lib, name, max_frames=None):
    g = np.load(GOLDEN / name)
    pcm = g["pcm"].astype(np.float32) / np.float32(32768)            # stored as int16
    frames, kbit = g["frames"], int(g["kbit"])
    S, F, C = pcm.shape[0], pcm.shape[1] // 1024, pcm.shape[2]
    if max_frames:
        F = min(F, max_frames)
        pcm = pcm[:, :F * 1024]
    enc = _at3_enc(lib, C, kbit)
    out = enc.encode(pcm, S)
    enc.close()
    assert out.shape == (S, F - 1, 1, frames.shape[-1])
    bad = np.argwhere((out[:, :, 0] != frames[:, :F - 1]).any(-1))
    assert bad.size == 0, f"first differing (stream, frame) = {bad[:4].tolist()}"


def check_at3_vs_oracle(lib, S, F, C, kbit=0, seed=1, kinds=("mix", "tones", "steps", None), **flags):
    """Bit-exact frames for S streams of mixed material against the reference."""
    streams = []
    for s in range(S):
        kind = kinds[s % len(kinds)]
        streams.append(tl.synth_streams(1, F, 1024, C, seed=seed + s)[0] if kind is None
                       else tl.synth_rich(F, 1024, C, seed=seed + s, kind=kind))
    pcm = np.stack(streams)
    enc = _at3_enc(lib, C, kbit, **flags)
    out = enc.encode(pcm, S)
    enc.close()
    checked = 0
    for s in range(S):
        want = oracle_at3(C, pcm[s].reshape(-1), kbit, **flags)
        bad = np.argwhere((out[s, :, 0] != want).any(-1))
        assert bad.size == 0, f"stream {s} ({kinds[s % len(kinds)]}): first differing frames {bad[:4, 0].tolist()}"
        checked += 1
    return checked


def check_at3_stage_taps(lib, C=2, F=12, kbit=0, seed=21, kind="mix"):
    """Intermediates against the reference's own sub-objects (needs oracle/_ref)."""
    tl.require_ref()
    pcm = tl.synth_rich(F, 1024, C, seed=seed, kind=kind)[None]
    recs, tracked, frames = tl.ref_at3_stages(C, pcm[0].reshape(-1), kbit)
    Fo = F - 1
    enc = _at3_enc(lib, C, kbit)
    enc.arm_taps()
    out = enc.encode(pcm, 1)
    if kbit != 64:          # bands are tapped before matrixing in the reference, after it here
        bands = enc.tap(ab.TAP_BANDS, (1, C, 4, 128 + 256 * F), np.float32)
        mine = bands[0][:, :, 128:128 + 256 * Fo].reshape(C, 4, Fo, 256).transpose(2, 0, 1, 3)
        assert np.array_equal(recs["bands"].view(np.uint32), mine.view(np.uint32))
    curves = enc.tap(ab.TAP_CURVES, (1, C, 4, Fo, 16), np.uint8)
    for f in range(Fo):
        for c in range(C):
            for b in range(3):
                n = int(recs[f, c]["n_points"][b])
                ref_pts = [tuple(int(v) for v in recs[f, c]["points"][b, i]) for i in range(n)]
                m = curves[0, c, b, f]
                assert ref_pts == [(int(m[1 + i]), int(m[8 + i])) for i in range(int(m[0]))], (f, c, b)
            assert int(recs[f, c]["n_points"][3]) == 0
    gs = enc.tap(ab.TAP_GSCALE, (1, Fo, C, 4, 4), np.float32)
    assert np.array_equal(recs["gscale"].view(np.uint32), gs[0][..., :3].view(np.uint32))
    chl = enc.tap(ab.TAP_CHLOUD, (1, Fo, C), np.float32)
    assert np.array_equal(recs["loud_term"].view(np.uint32), chl[0].view(np.uint32))
    loud = enc.tap(ab.TAP_LOUDNESS, (1, Fo), np.float32)
    assert np.array_equal(tracked.view(np.uint32), loud[0].view(np.uint32))
    sfi = enc.tap(ab.TAP_SFI, (1, Fo, C, 32), np.uint8)
    assert np.array_equal(recs["sfi"], sfi[0].astype(np.int32))
    en = enc.tap(ab.TAP_ENERGY, (1, Fo, C, 32), np.float32)
    assert np.array_equal(recs["energy"].view(np.uint32), en[0].view(np.uint32))
    enc.close()
    assert np.array_equal(out[0, :, 0], frames)
    return True


def check_at3_main_loop(lib, C=2, kbit=0, seconds=0.4):
    """Through src/main.cpp's PCM pump (TPCMEngine(4096) + look-ahead): the frames the reference CLI
    would hand to WriteFrame for a WAV of `seconds`, against ours on the frames the lambda received."""
    tl.require_ref()
    n = int(44100 * seconds)
    pcm = tl.synth_rich((n + 1023) // 1024, 1024, C, seed=31)[:n]
    payload, sizes = tl.ref_encode(3, C, pcm.reshape(-1), total=n, bitrate_kbit=kbit)
    fs = int(sizes[0])
    assert (sizes == fs).all()
    want = payload.reshape(-1, fs)
    view = tl.engine_view(pcm, C, 1024, total=n, lookahead=1)
    enc = _at3_enc(lib, C, kbit)
    out = enc.encode(view, 1)
    enc.close()
    assert out.shape[1] == want.shape[0]
    assert np.array_equal(out[0, :, 0], want)
    return True


def check_at3_batch_split_invariance(lib, S=2, F=13, C=2, kbit=0, cuts=(4, 5)):
    """Streams continue across calls: [0,a) + [a,b) + [b,F) must equal one batch of F frames
    (carried look-ahead frame, MDCT half, PrevOverlapGainScale, CurveCtx, loudness)."""
    pcm = np.stack([tl.synth_rich(F, 1024, C, seed=40 + s, kind=("steps", "mix")[s % 2]) for s in range(S)])
    enc = _at3_enc(lib, C, kbit)
    whole = enc.encode(pcm, S)
    enc.reset()
    a, b = cuts[0], cuts[0] + cuts[1]
    parts = [enc.encode(pcm[:, :a * 1024], S), enc.encode(pcm[:, a * 1024:b * 1024], S), enc.encode(pcm[:, b * 1024:], S)]
    enc.close()
    assert [p.shape[1] for p in parts] == [a - 1, cuts[1], F - b]
    assert np.array_equal(np.concatenate(parts, axis=1), whole)


def check_at3_stream_independence(lib, C=2, F=7, kbit=0):
    pcm = np.stack([tl.synth_rich(F, 1024, C, seed=50 + s, kind=("mix", "tones", "steps")[s % 3]) for s in range(4)])
    enc = _at3_enc(lib, C, kbit)
    batch = enc.encode(pcm, 4)
    enc.reset()
    rev = enc.encode(pcm[::-1].copy(), 4)
    enc.reset()
    one = enc.encode(pcm[2:3].copy(), 1)
    enc.close()
    assert np.array_equal(batch, rev[::-1])
    assert np.array_equal(batch[2:3], one)


def check_at3_edge_inputs(lib, kbit=0):
    """Silence, digital full scale DC, Nyquist, a single impulse; a one-frame first batch (no output)."""
    F = 6
    n = F * 1024
    cases = {
        "silence": np.zeros((n, 2), np.float32),
        "full_scale_dc": np.full((n, 2), 32767 / 32768, np.float32),
        "nyquist": np.tile(np.array([[1.0], [-1.0]], np.float32), (n // 2, 2)) * np.float32(32767 / 32768),
        "impulse": np.zeros((n, 2), np.float32),
    }
    cases["impulse"][1500, 0] = -1.0
    checked = 0
    for name, x in cases.items():
        enc = _at3_enc(lib, 2, kbit)
        out = enc.encode(x, 1)
        enc.close()
        want = oracle_at3(2, x.reshape(-1), kbit)
        assert np.array_equal(out[0, :, 0], want), name
        checked += 1
    enc = _at3_enc(lib, 2, kbit)
    first = enc.encode(cases["impulse"][:1024], 1)
    assert first.shape[1] == 0
    rest = enc.encode(cases["impulse"][1024:], 1)
    enc.close()
    want = oracle_at3(2, cases["impulse"].reshape(-1), kbit)
    assert np.array_equal(rest[0, :, 0], want)
    return checked


def check_at3_errors(lib):
    import pytest
    with pytest.raises(ab.AtdeError):
        ab.Encoder(ab.CODEC_ATRAC3, 2, bitrate=400 * 1024, lib=lib)      # beyond the largest container
    with pytest.raises(ab.AtdeError):
        ab.Encoder(2, 2, lib=lib)                                        # no such codec
    enc = ab.Encoder(ab.CODEC_ATRAC3, 2, lib=lib)
    assert (enc.frame_samples, enc.units_per_frame, enc.unit_bytes, enc.lookahead) == (1024, 1, 384, 1)
    enc.close()
    enc = ab.Encoder(ab.CODEC_ATRAC3, 2, bitrate=64 * 1024, lib=lib)
    assert enc.unit_bytes == 192
    enc.close()


# ---------------------------------------------------------------------------------------------
# ATRAC3plus, stage by stage (the GHA stage is not built yet: atracdenc_b200/csrc/at3p_stage_api.h)
def _at3p_signal(S, F, C, seed):
    kinds = ("mix", "tones", "steps", None)
    streams = []
    for s in range(S):
        kind = kinds[s % len(kinds)]
        streams.append(tl.synth_streams(1, F, 2048, C, seed=seed + s)[0] if kind is None
                       else tl.synth_rich(F, 2048, C, seed=seed + s, kind=kind))
    return np.stack(streams)                                       # [S][F*2048][C]


def at3p_stage_pqf(lib, pcm, S, C, F):
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    out = np.zeros((S, C, F, 2048), np.float32)
    rc = lib.atde_at3p_stage_pqf(pcm.ctypes.data_as(tl.P), S, C, F, out.ctypes.data_as(tl.P))
    assert rc == 0, f"atde_at3p_stage_pqf -> {rc}"
    return out


def at3p_stage_mdct(lib, resid, S, C, F):
    resid = np.ascontiguousarray(resid, dtype=np.float32)
    out = np.zeros((S, F, C, 2048), np.float32)
    rc = lib.atde_at3p_stage_mdct(resid.ctypes.data_as(tl.P), S, C, F, out.ctypes.data_as(tl.P))
    assert rc == 0, f"atde_at3p_stage_mdct -> {rc}"
    return out


def check_at3p_pqf(lib, S=3, F=6, C=2, seed=900):
    """PQF analysis bit-exact against at3plus_pqf_do_analyse (fresh context per stream and channel)."""
    pcm = _at3p_signal(S, F, C, seed)
    got = at3p_stage_pqf(lib, pcm, S, C, F)
    tl.require_ref()
    for s in range(S):
        for c in range(C):
            want = tl.ref_at3p_pqf(pcm[s, :, c])
            bad = np.argwhere(got[s, c].view(np.uint32) != want.view(np.uint32))
            assert bad.size == 0, f"stream {s} ch {c}: first differing (frame, index) = {bad[:4].tolist()}"
    return S


def check_at3p_mdct(lib, S=2, F=6, C=2, seed=910):
    """MDCT bit-exact against TAt3pMDCT::Do fed with the residual the reference encoder produced
    (work buffer after the tone filter, scaled like at3p.cpp:150-153)."""
    tl.require_ref()
    pcm = _at3p_signal(S, F + 1, C, seed)
    resid = np.zeros((S, C, F, 2048), np.float32)
    want = np.zeros((S, F, C, 2048), np.float32)
    for s in range(S):
        st = tl.ref_at3p_stages(C, pcm[s].reshape(-1))
        assert st["n"] == F
        # tmp[i] = x[i] / (32768.0 / 1.122018): double division, rounded to float on the store
        r = (st["work_out"].astype(np.float64) / (32768.0 / 1.122018)).astype(np.float32)
        resid[s] = r.transpose(1, 0, 2)
        want[s] = st["specs"]
        for c in range(C):                                           # the harness' own MDCT entry agrees with the encoder's
            assert np.array_equal(tl.ref_at3p_mdct(r[:, c]).view(np.uint32), st["specs"][:, c].view(np.uint32))
    got = at3p_stage_mdct(lib, resid, S, C, F)
    bad = np.argwhere(got.view(np.uint32) != want.view(np.uint32))
    assert bad.size == 0, f"first differing (stream, frame, ch, line) = {bad[:4].tolist()}"
    return S


def at3p_stage_pack(lib, specs, tones, C):
    specs = np.ascontiguousarray(specs, dtype=np.float32)
    tones = np.ascontiguousarray(tones)
    assert lib.atde_at3p_tone_block_size() == tl.AT3P_GHA_REC.itemsize
    U = specs.shape[0]
    out = np.zeros((U, 2048), np.uint8)
    rc = lib.atde_at3p_stage_pack(specs.ctypes.data_as(tl.P), tones.ctypes.data_as(tl.P), U, C, out.ctypes.data_as(tl.P))
    assert rc == 0, f"atde_at3p_stage_pack -> {rc}"
    return out


def check_at3p_pack(lib, S=2, F=6, C=2, seed=920, loud=False):
    """Frame packer (scale, quantise, code-table choice, tonal block, bit writer) bit-exact against
    TAt3PBitStream::WriteFrame, fed with the spectra and the tone data of the reference encoder: the
    frame written in call t carries the GHA result of call t-1 (`delay`, at3p.cpp:127-131,186-190)."""
    tl.require_ref()
    pcm = _at3p_signal(S, F + 1, C, seed)
    if loud:
        pcm = tl.quantise(np.clip(pcm.astype(np.float64) * 12.0, -1.0, 1.0))       # forces the unit-dropping loop
    for s in range(S):
        st = tl.ref_at3p_stages(C, pcm[s].reshape(-1))
        assert st["n"] == F
        tones = np.zeros(F, tl.AT3P_GHA_REC)
        tones[1:] = st["gha"][:-1]
        got = at3p_stage_pack(lib, st["specs"], tones, C)
        bad = np.argwhere((got != st["frames"]).any(-1))
        assert bad.size == 0, f"stream {s}: first differing frames {bad[:4].ravel().tolist()}, " \
                              f"first byte {np.argwhere(got[bad[0, 0]] != st['frames'][bad[0, 0]])[:3].ravel().tolist()}"
    return S


def check_at3p_pack_random(lib, U=12, C=2, seed=940):
    """Packer on random spectra of growing level: walks TTonalComponentEncoder's unit-dropping loop well
    below 28 units, which no natural signal reaches; tone data borrowed from a real encode."""
    tl.require_ref()
    rng = np.random.default_rng(seed)
    st = tl.ref_at3p_stages(C, _at3p_signal(1, 5, C, seed)[0].reshape(-1))
    tones = np.zeros(U, tl.AT3P_GHA_REC)
    specs = np.zeros((U, C, 2048), np.float32)
    for u in range(U):
        level = 10.0 ** rng.uniform(-4, 0)
        specs[u] = rng.uniform(-level, level, (C, 2048)) * (rng.uniform(0, 1, (C, 2048)) < rng.uniform(0.05, 1.0))
        if u % 3:
            tones[u] = st["gha"][u % st["n"]]
    specs[U - 1] = 0.0
    specs[0, 0, 5] = 1.0                                              # == MAX_SCALE: last table entry, value clipped to 0.99999
    want = tl.ref_at3p_pack(C, specs, tones)
    got = at3p_stage_pack(lib, specs, tones, C)
    nq = sorted({int(f[0] & 0x1f) + 1 for f in want})
    bad = np.argwhere((got != want).any(-1))
    assert bad.size == 0, f"differing frames {bad[:4].ravel().tolist()} (unit counts seen: {nq})"
    return nq


def at3p_stage_tone_filter(lib, bands, old, now, nxt, C):
    bands = np.ascontiguousarray(bands, dtype=np.float32)
    U = bands.shape[0]
    out = np.zeros((U, C, 2048), np.float32)
    a = [np.ascontiguousarray(t) for t in (old, now, nxt)]
    rc = lib.atde_at3p_stage_tone_filter(bands.ctypes.data_as(tl.P), a[0].ctypes.data_as(tl.P), a[1].ctypes.data_as(tl.P),
                                         a[2].ctypes.data_as(tl.P), U, C, out.ctypes.data_as(tl.P))
    assert rc == 0, f"atde_at3p_stage_tone_filter -> {rc}"
    return out


def _shift(recs, k):
    out = np.zeros_like(recs)
    if k < len(recs):
        out[k:] = recs[:len(recs) - k]
    return out


def check_at3p_tone_filter(lib, S=2, F=8, C=2, seed=960):
    """Tone subtraction + MDCT input scaling bit-exact against the work buffer TGhaProcessorBase::ApplyFilter
    leaves behind, fed with the reference's own GHA results of three consecutive calls."""
    tl.require_ref()
    pcm = _at3p_signal(S, F + 1, C, seed)
    for s in range(S):
        st = tl.ref_at3p_stages(C, pcm[s].reshape(-1))
        want = (st["work_out"].astype(np.float64) / (32768.0 / 1.122018)).astype(np.float32)
        got = at3p_stage_tone_filter(lib, st["work_in"], _shift(st["gha"], 2), _shift(st["gha"], 1), st["gha"], C)
        bad = np.argwhere(got.view(np.uint32) != want.view(np.uint32))
        assert bad.size == 0, f"stream {s}: first differing (frame, ch, sample) = {bad[:4].tolist()}"
        assert np.abs(st["work_out"] - st["work_in"]).max() > 1.0          # tones were actually subtracted
    return S


def check_at3p_chain_after_gha(lib, S=2, F=8, C=2, seed=970):
    """PCM -> PQF -> [reference GHA results] -> tone filter -> MDCT -> packer == the reference's frames:
    everything of the ATRAC3plus path except the tone search itself, chained on the device kernels."""
    tl.require_ref()
    pcm = _at3p_signal(S, F + 1, C, seed)
    bands = at3p_stage_pqf(lib, pcm, S, C, F + 1)                        # [S][C][F+1][2048]
    for s in range(S):
        st = tl.ref_at3p_stages(C, pcm[s].reshape(-1))
        assert st["n"] == F
        # output o encodes PQF frame o-1 (zeros for o = 0): at3p.cpp:113-121,181-185
        work = np.zeros((F, C, 2048), np.float32)
        work[1:] = bands[s].transpose(1, 0, 2)[:F - 1]
        assert np.array_equal(work, st["work_in"])
        resid = at3p_stage_tone_filter(lib, work, _shift(st["gha"], 2), _shift(st["gha"], 1), st["gha"], C)
        specs = at3p_stage_mdct(lib, resid.transpose(1, 0, 2)[None], 1, C, F)[0]      # [F][C][2048]
        frames = at3p_stage_pack(lib, specs, _shift(st["gha"], 1), C)
        bad = np.argwhere((frames != st["frames"]).any(-1))
        assert bad.size == 0, f"stream {s}: differing frames {bad[:4].ravel().tolist()}"
    return S


def check_trig_replicas(lib, n=2000000, seed=77):
    """Device replicas of glibc sin / cos / atan / sincosf against the live libm of the box, on the argument
    ranges the tone search visits (phases up to ~2^9, ratios of any magnitude) and beyond."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    rng = np.random.default_rng(seed)
    xs = np.concatenate([
        rng.uniform(-4, 4, n // 4), rng.uniform(0, 420, n // 4).astype(np.float32).astype(np.float64),
        np.ldexp(rng.uniform(0.5, 1, n // 4), rng.integers(-40, 26, n // 4)) * rng.choice([-1, 1], n // 4),
        rng.uniform(-1e8, 1e8, n // 4), np.array([0.0, -0.0, 0.126, 0.855469, 2.426265, 1e-300, np.pi, np.pi / 2]),
    ])
    xs = np.ascontiguousarray(xs[np.abs(xs) < 105414000.0])
    def dev(fn, x):
        y = np.empty_like(x)
        rc = lib.atde_at3p_debug_trig(fn, x.ctypes.data_as(tl.P), y.ctypes.data_as(tl.P), ctypes.c_longlong(x.size))
        assert rc == 0
        return y
    for fn, name in ((0, "sin"), (1, "cos"), (2, "atan")):
        f = getattr(libm, name)
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_double]
        x = xs if name != "atan" else np.concatenate([xs, np.ldexp(rng.uniform(0.5, 1, n // 4), rng.integers(-60, 70, n // 4))])
        x = np.ascontiguousarray(x[:: max(1, x.size // 60000)])            # the ctypes loop is the slow part
        want = np.array([f(float(v)) for v in x])
        got = dev(fn, x)
        bad = np.argwhere(want.view(np.uint64) != got.view(np.uint64))
        assert bad.size == 0, (name, x[bad[:3].ravel()], want[bad[:3].ravel()], got[bad[:3].ravel()])
    # sincosf: numpy has no glibc sinf; go through libm.sincosf
    sc = libm.sincosf
    sc.restype = None
    sc.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    xf = np.concatenate([rng.uniform(-420, 420, 20000), np.ldexp(rng.uniform(0.5, 1, 20000), rng.integers(-30, 60, 20000))]).astype(np.float32)
    s, c = ctypes.c_float(), ctypes.c_float()
    ws, wc = np.empty(xf.size, np.float32), np.empty(xf.size, np.float32)
    for i, v in enumerate(xf):
        sc(float(v), ctypes.byref(s), ctypes.byref(c))
        ws[i], wc[i] = s.value, c.value
    xd = np.ascontiguousarray(xf.astype(np.float64))
    assert np.array_equal(dev(3, xd).astype(np.float32).view(np.uint32), ws.view(np.uint32))
    assert np.array_equal(dev(4, xd).astype(np.float32).view(np.uint32), wc.view(np.uint32))


def at3p_stage_gha(lib, bands, S, C, F):
    bands = np.ascontiguousarray(bands, dtype=np.float32)
    assert lib.atde_at3p_tone_block_size() == tl.AT3P_GHA_REC.itemsize
    out = np.zeros((S, F), tl.AT3P_GHA_REC)
    rc = lib.atde_at3p_stage_gha(bands.ctypes.data_as(tl.P), S, C, F, out.ctypes.data_as(tl.P))
    assert rc == 0, f"atde_at3p_stage_gha -> {rc}"
    return out


def _tone_records_equal(got, want, C):
    """Compares the meaningful fields of two flattened TAt3PGhaData records (the reference leaves stale
    ToneSharing flags / wave slots beyond NumToneBands behind)."""
    if int(got["present"]) != int(want["present"]):
        return "present"
    if not want["present"]:
        return None
    for k in ("num_tone_bands", "second_is_leader"):
        if int(got[k]) != int(want[k]):
            return k
    ntb = int(want["num_tone_bands"])
    if C == 2 and not np.array_equal(got["tone_sharing"][:ntb] != 0, want["tone_sharing"][:ntb] != 0):
        return "tone_sharing"
    for ch in range(C):
        if int(got["n_sb"][ch]) != int(want["n_sb"][ch]):
            return f"n_sb[{ch}]"
        for sb in range(ntb):
            if ch == 1 and want["tone_sharing"][sb]:
                continue
            g, w = got["sb"][ch][sb], want["sb"][ch][sb]
            if int(g[1]) != int(w[1]) or (int(w[1]) and int(g[0]) != int(w[0])) or (int(w[1]) and not np.array_equal(g[2:], w[2:])):
                return f"sb[{ch}][{sb}] {g.tolist()} vs {w.tolist()}"
            n, i0 = int(w[1]), int(w[0])
            if not np.array_equal(got["params"][ch][i0:i0 + n], want["params"][ch][i0:i0 + n]):
                return f"params[{ch}] sb {sb}: {got['params'][ch][i0:i0 + n].tolist()} vs {want['params'][ch][i0:i0 + n].tolist()}"
    return None


def check_at3p_gha(lib, S=2, F=4, C=2, seed=980):
    """The tone search (DoAnalize without the filter) against the reference's GHA results, frame by frame."""
    tl.require_ref()
    pcm = _at3p_signal(S, F + 1, C, seed)
    bands = at3p_stage_pqf(lib, pcm, S, C, F + 1)
    got = at3p_stage_gha(lib, bands, S, C, F + 1)
    for s in range(S):
        st = tl.ref_at3p_stages(C, pcm[s].reshape(-1))
        assert st["n"] == F
        for f in range(F):
            why = _tone_records_equal(got[s, f], st["gha"][f], C)
            assert why is None, f"stream {s} frame {f}: {why}"
    return S


def check_at3p_full_chain(lib, S=2, F=6, C=2, seed=990):
    """PCM -> PQF -> tone search -> tone filter -> MDCT -> packer, every stage on the device kernels,
    == the frames of the reference encoder (fresh streams)."""
    tl.require_ref()
    pcm = _at3p_signal(S, F + 1, C, seed)
    bands = at3p_stage_pqf(lib, pcm, S, C, F + 1)                        # [S][C][F+1][2048]
    tones = at3p_stage_gha(lib, bands, S, C, F + 1)                      # [S][F+1]; call o+1 analyses frame o
    for s in range(S):
        st = tl.ref_at3p_stages(C, pcm[s].reshape(-1))
        assert st["n"] == F
        gha = tones[s, :F]
        work = np.zeros((F, C, 2048), np.float32)                        # output o encodes PQF frame o-1
        work[1:] = bands[s].transpose(1, 0, 2)[:F - 1]
        resid = at3p_stage_tone_filter(lib, work, _shift(gha, 2), _shift(gha, 1), gha, C)
        specs = at3p_stage_mdct(lib, resid.transpose(1, 0, 2)[None], 1, C, F)[0]
        frames = at3p_stage_pack(lib, specs, _shift(gha, 1), C)
        bad = np.argwhere((frames != st["frames"]).any(-1))
        assert bad.size == 0, f"stream {s}: differing frames {bad[:4].ravel().tolist()}"
    return S


# ---------------------------------------------------------------------------------------------
# ATRAC3plus through the C ABI (atde_create(ATDE_CODEC_ATRAC3PLUS))
def check_at3p_vs_oracle(lib, S=3, F=7, C=2, seed=1200):
    """Whole ATRAC3plus encoder through atde_encode_batch == the reference encoder's frames."""
    pcm = _at3p_signal(S, F, C, seed)
    enc = ab.Encoder(ab.CODEC_ATRAC3PLUS, C, lib=lib)
    assert (enc.frame_samples, enc.units_per_frame, enc.unit_bytes, enc.lookahead) == (2048, 1, 2048, 1)
    out = enc.encode(pcm, S)
    enc.close()
    assert out.shape == (S, F - 1, 1, 2048)
    tl.require_ref()
    for s in range(S):
        st = tl.ref_at3p_stages(C, pcm[s].reshape(-1))
        assert st["n"] == F - 1
        bad = np.argwhere((out[s, :, 0] != st["frames"]).any(-1))
        assert bad.size == 0, f"stream {s}: differing frames {bad[:4].ravel().tolist()}"
    return S


def check_at3p_batch_split_invariance(lib, S=2, F=9, C=2, cuts=(1, 3, 2), seed=1300):
    """Streams continue across calls (carried PQF history, two PQF frames, MDCT overlap, the last two GHA
    results, envelope history): the pieces must equal one batch."""
    pcm = _at3p_signal(S, F, C, seed)
    enc = ab.Encoder(ab.CODEC_ATRAC3PLUS, C, lib=lib)
    whole = enc.encode(pcm, S)
    enc.reset()
    parts, pos = [], 0
    for n in list(cuts) + [F - sum(cuts)]:
        parts.append(enc.encode(pcm[:, pos * 2048:(pos + n) * 2048], S))
        pos += n
    enc.close()
    assert [p.shape[1] for p in parts] == [cuts[0] - 1] + list(cuts[1:]) + [F - sum(cuts)]
    assert np.array_equal(np.concatenate(parts, axis=1), whole)


def check_host_chunking(lib, codec, C=2, S=11, F=5, seed=1700, variants=(None, "4", "2", "1", "2/i16")):
    """atde_encode_batch() splits the streams of a batch into chunks (a short first one, then equal ones, on three
    pipeline slots; ATRAC3 PCM through a ring of four staging buffers).  Forced down to four streams per chunk
    (1 + 4 + 4 + ...), to two and to one (the staging ring wraps) — or, with S >= 27, to eight (2 + 7 + 7 + 5 + 4 + 2: a
    tapered tail) — the result must equal the single-chunk batch, also on the continuation batch that starts from
    carried state, and the int16 entry point must agree.  Variant "/i16": the library's own plan through the int16 entry point."""
    import os
    step = {ab.CODEC_ATRAC1: 512, ab.CODEC_ATRAC3: 1024, ab.CODEC_ATRAC3PLUS: 2048}[codec]
    rng = np.random.default_rng(seed)
    pcm = np.stack([tl.synth_rich(2 * F, step, C, seed=seed + s, kind=("mix", "tones", "steps")[s % 3]) for s in range(S)])
    pcm = (pcm * rng.uniform(0.2, 1.0, size=(S, 1, 1))).astype(np.float32)
    old = os.environ.pop("ATDE_CHUNK_STREAMS", None)
    try:
        outs = []
        q = np.clip(np.rint(pcm * 32768.0), -32768, 32767).astype(np.int16)
        pcm = (q.astype(np.float32) * np.float32(1.0 / 32768.0)).astype(np.float32)
        for streams in variants:
            os.environ.pop("ATDE_CHUNK_STREAMS", None)
            if streams and streams.split("/")[0]:                  # ("/i16": the library's own plan, int16 entry point)
                os.environ["ATDE_CHUNK_STREAMS"] = streams.split("/")[0]
            enc = ab.Encoder(codec, C, lib=lib)
            if streams and streams.endswith("i16"):
                a = enc.encode_i16(q[:, :F * step], S, want_sizes=True)
                b = enc.encode_i16(q[:, F * step:], S, want_sizes=True)
            else:
                a = enc.encode(pcm[:, :F * step], S, want_sizes=True)
                b = enc.encode(pcm[:, F * step:], S, want_sizes=True)
            enc.close()
            outs.append((a, b))
        for other in outs[1:]:
            for (x, xs), (y, ys) in zip(outs[0], other):
                assert np.array_equal(xs, ys)
                assert np.array_equal(x, y)
    finally:
        os.environ.pop("ATDE_CHUNK_STREAMS", None)
        if old is not None:
            os.environ["ATDE_CHUNK_STREAMS"] = old


def check_i16_ingest(lib, codec, C=2, S=3, F=5, seed=1800):
    """atde_encode_batch_i16 == atde_encode_batch on the floats the reference's reader would produce
    (libsndfile: int16 * (1 / 0x8000)), across a continuation batch; extreme sample values included."""
    step = {ab.CODEC_ATRAC1: 512, ab.CODEC_ATRAC3: 1024, ab.CODEC_ATRAC3PLUS: 2048}[codec]
    pcm = np.stack([tl.synth_rich(2 * F, step, C, seed=seed + s, kind=("mix", "tones", "steps")[s % 3]) for s in range(S)])
    q = np.clip(np.rint(pcm * 32768.0), -32768, 32767).astype(np.int16)
    q[0, :4, :] = np.array([[-32768, 32767], [32767, -32768], [0, -1], [1, 0]], np.int16)[:, :C]
    f = (q.astype(np.float32) * np.float32(1.0 / 32768.0)).astype(np.float32)
    enc = ab.Encoder(codec, C, lib=lib)
    a1, s1 = enc.encode(f[:, :F * step], S, want_sizes=True)
    a2, s2 = enc.encode(f[:, F * step:], S, want_sizes=True)
    enc.reset()
    b1, t1 = enc.encode_i16(q[:, :F * step], S, want_sizes=True)
    b2, t2 = enc.encode_i16(q[:, F * step:], S, want_sizes=True)
    enc.close()
    assert np.array_equal(s1, t1) and np.array_equal(s2, t2)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)


def check_at3p_gha_masks(lib, masks=(0, 1, 2, 3, 4, 5, 6), S=2, F=6, C=2, seed=1900):
    """TAt3PEnc::TSettings::UseGha other than GHA_ENABLED (`--advanced ghadbg=N`, at3p.cpp:143-177): without
    PASS_INPUT the tones are subtracted from zeros, without WRITE_RESIUDAL the MDCT sees zeros, without WRITE_TONAL
    no tone block is written.  Frames against the reference run with the same mask, across two batches."""
    import pytest
    pcm = _at3p_signal(S, F, C, seed)
    with pytest.raises(ab.AtdeError):
        ab.Encoder(ab.CODEC_ATRAC3PLUS, C, lib=lib, gha_flags=8 | 7)          # GHA_WIDEBAND: not built
    tl.require_ref()
    for mask in masks:
        enc = ab.Encoder(ab.CODEC_ATRAC3PLUS, C, lib=lib, gha_flags=mask)
        a = enc.encode(pcm[:, :3 * 2048], S)
        b = enc.encode(pcm[:, 3 * 2048:], S)
        enc.close()
        out = np.concatenate([a, b], axis=1)
        for s in range(S):
            st = tl.ref_at3p_stages(C, pcm[s].reshape(-1), gha_flags=mask)
            assert st["n"] == F - 1
            bad = np.argwhere((out[s, :, 0] != st["frames"]).any(-1))
            assert bad.size == 0, f"mask {mask} stream {s}: differing frames {bad[:4].ravel().tolist()}"
    return len(masks)


def check_at3p_edge_inputs(lib, C=2, F=5):
    """Silence, full-scale DC, Nyquist, one impulse, a loud pure sine, two sines either side of a PQF subband
    edge, a sine that stops mid-frame; and a one-frame first batch (no output) followed by the rest."""
    n = F * 2048
    t = np.arange(n, dtype=np.float64)
    def chans(x):
        x = np.asarray(x, np.float32)
        return np.stack([x, (x * np.float32(0.5) if C == 2 else x)][:C], axis=1) if x.ndim == 1 else x
    sine = lambda f, a=0.5: (a * np.sin(2 * np.pi * f / 44100.0 * t))
    stop = sine(1000.0, 0.7); stop[2 * 2048 + 700:] = 0.0
    cases = {
        "silence": np.zeros((n, C), np.float32),
        "full_scale_dc": np.full((n, C), 32767 / 32768, np.float32),
        "nyquist": chans(np.tile(np.array([1.0, -1.0]), n // 2) * (32767 / 32768)),
        "impulse": np.zeros((n, C), np.float32),
        "sine_1k": chans(sine(1000.0, 0.9)),
        "sines_at_subband_edge": chans(sine(1378.125 - 20.0, 0.4) + sine(1378.125 + 20.0, 0.4)),
        "sine_stops": chans(stop),
    }
    cases["impulse"][3000, 0] = -1.0
    checked = 0
    tl.require_ref()
    for name, x in cases.items():
        x = tl.quantise(np.ascontiguousarray(x, np.float32))
        enc = ab.Encoder(ab.CODEC_ATRAC3PLUS, C, lib=lib)
        out = enc.encode(x, 1)
        enc.close()
        assert out.shape == (1, F - 1, 1, 2048), name
        st = tl.ref_at3p_stages(C, x.reshape(-1))
        bad = np.argwhere((out[0, :, 0] != st["frames"]).any(-1))
        assert bad.size == 0, f"{name}: differing frames {bad[:4].ravel().tolist()}"
        checked += 1
    x = tl.quantise(np.ascontiguousarray(cases["sine_stops"], np.float32))
    enc = ab.Encoder(ab.CODEC_ATRAC3PLUS, C, lib=lib)
    first = enc.encode(x[:2048], 1)
    assert first.shape[1] == 0
    rest = enc.encode(x[2048:], 1)
    enc.close()
    assert np.array_equal(rest[0, :, 0], tl.ref_at3p_stages(C, x.reshape(-1))["frames"])
    return checked


def check_at3p_stream_independence(lib, C=2, F=5):
    pcm = _at3p_signal(4, F, C, 2100)
    enc = ab.Encoder(ab.CODEC_ATRAC3PLUS, C, lib=lib)
    batch = enc.encode(pcm, 4)
    enc.reset()
    rev = enc.encode(pcm[::-1].copy(), 4)
    enc.reset()
    one = enc.encode(pcm[2:3].copy(), 1)
    enc.close()
    assert np.array_equal(batch, rev[::-1])
    assert np.array_equal(batch[2:3], one)


def check_at3p_golden(lib, name, max_frames=None):
    """Committed reference output (tests/golden/make_golden.py): works on a box without oracle/_ref."""
    g = np.load(GOLDEN / name)
    pcm = g["pcm"].astype(np.float32) / np.float32(32768)            # stored as int16
    frames = g["frames"]
    S, F, C = pcm.shape[0], pcm.shape[1] // 2048, pcm.shape[2]
    if max_frames:
        F = min(F, max_frames)
        pcm = pcm[:, :F * 2048]
    enc = ab.Encoder(ab.CODEC_ATRAC3PLUS, C, lib=lib)
    out = enc.encode(pcm, S)
    enc.close()
    assert out.shape == (S, F - 1, 1, 2048)
    bad = np.argwhere((out[:, :, 0] != frames[:, :F - 1]).any(-1))
    assert bad.size == 0, f"first differing (stream, frame) = {bad[:4].tolist()}"


def check_at3_gain_trace_taps(lib, S=2, F=6, C=2, kbit=0, seed=2100):
    """atde_set_gain_trace (the data behind `--yaml-log`): the trace instance of the gain kernel analyses all four bands;
    on bands 0..2 its envelope must be bit-equal to the encode path's, its high-frequency ratio may differ from the
    encode path's tree sum only in the last place, `next_level` is a finite RMS, band 3 is analysed too — and the frames
    are the frames of an encoder without the trace."""
    pcm = np.stack([tl.synth_rich(F, 1024, C, seed=seed + s, kind=("mix", "steps")[s % 2]) for s in range(S)]).astype(np.float32)
    plain = ab.Encoder(ab.CODEC_ATRAC3, C, bitrate=kbit * 1024, lib=lib)
    want = plain.encode(pcm, S)
    plain.close()
    enc = ab.Encoder(ab.CODEC_ATRAC3, C, bitrate=kbit * 1024, lib=lib)
    enc.set_gain_trace(True)
    got = enc.encode(pcm, S)
    n_out = F - 1
    gain = enc.tap(ab.TAP_GAIN, (S, C, 3, n_out, 96), np.float32)
    tgain = enc.tap(ab.TAP_TRACE_GAIN, (S, C, 4, n_out, 96), np.float32)
    tstat = enc.tap(ab.TAP_TRACE_STAT, (S, C, 4, n_out, 4), np.float32)
    enc.close()
    assert np.array_equal(got, want), "the trace must not change the encoded frames"
    assert np.array_equal(tgain[:, :, :3].view(np.uint32), gain.view(np.uint32)), "trace envelope != encode path's envelope"
    assert np.isfinite(tstat).all() and (tstat[..., 3] >= 0).all() and (tstat[..., 0] >= 0).all() and (tstat[..., 0] <= 1.0001).all()
    assert np.abs(tgain[:, :, 3]).sum() > 0, "band 3 was not analysed"
    # curHpfEnergy is the mean of the 32 sub-frame levels, in the reference's order
    mean = np.zeros(tgain.shape[:4], np.float32)
    for i in range(32):
        mean = (mean + tgain[..., i]).astype(np.float32)
    mean = (mean / np.float32(32.0)).astype(np.float32)
    assert np.array_equal(mean.view(np.uint32), tstat[..., 1].view(np.uint32))
