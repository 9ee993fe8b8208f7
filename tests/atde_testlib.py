"""Shared helpers for the test-suite: oracle bindings (CHECKERS ONLY) and deterministic signals."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_SO = ROOT / "oracle" / "_ref" / "libatde_ref.so"
PORT_SO = ROOT / "oracle" / "_build" / "libatde_oracle.so"
EMU_SO = ROOT / "tests" / "cpuemu" / "_build" / "libatde_emu.so"
P = ctypes.c_void_p


def build_port():
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "port"])
    return PORT_SO


def build_emu():
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "tests" / "cpuemu")])
    return EMU_SO


_ref = None
_port = None


def ref_lib():
    """The UNMODIFIED reference encoder compiled into oracle/_ref (prebuilt on the authoring box)."""
    global _ref
    if _ref is None:
        if not REF_SO.exists():
            if Path("/root/reference/src").exists():
                subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "ref"])
            else:
                return None
        _ref = ctypes.CDLL(str(REF_SO))
        _ref.ref_encode.restype = ctypes.c_long
        for f in ("ref_log10f", "ref_log2f"):
            getattr(_ref, f).restype = ctypes.c_float
            getattr(_ref, f).argtypes = [ctypes.c_float]
    return _ref


def require_ref(what="this check"):
    """Parity against the reference needs oracle/_ref.  Without it the calling test is SKIPPED with a reason
    (never a pass); under ATDE_REQUIRE_REF=1 — set by tests/conftest.py on a box with a GPU, where the prebuilt
    oracle/_ref must have travelled with the snapshot — it is a hard failure."""
    lib = ref_lib()
    if lib is not None:
        return lib
    import os
    msg = f"oracle/_ref/libatde_ref.so is missing: {what} cannot be compared with the reference"
    if os.environ.get("ATDE_REQUIRE_REF") == "1":
        raise AssertionError(msg + " (ATDE_REQUIRE_REF=1)")
    import pytest
    pytest.skip(msg)


def port_lib():
    """The plain-C restatement in oracle/*.c."""
    global _port
    if _port is None:
        build_port()
        _port = ctypes.CDLL(str(PORT_SO))
        for f in ("og_logf", "og_log10f", "og_log2f"):
            getattr(_port, f).restype = ctypes.c_float
            getattr(_port, f).argtypes = [ctypes.c_float]
    return _port


def quantise(x):
    """int16-quantise then /32768 -> what libsndfile hands the reference for a 16-bit WAV."""
    return (np.rint(np.asarray(x, dtype=np.float64) * 32767).clip(-32768, 32767).astype(np.int16)
            .astype(np.float32) / np.float32(32768))


def ref_encode(codec, channels, pcm, total=None, bitrate_kbit=0, no_gain=0, no_tonal=0, bfu=0,
               window_auto=1, window_mask=0):
    """Runs src/main.cpp's PCM loop over `pcm` (interleaved).  Returns (payload bytes, sizes)."""
    lib = ref_lib()
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    n = pcm.size // channels
    total = n if total is None else total
    out = np.zeros(max(1 << 20, pcm.size), dtype=np.uint8)
    sizes = np.zeros(max(1 << 12, n // 256 + 64), dtype=np.int32)
    nb = ctypes.c_long()
    k = lib.ref_encode(codec, channels, pcm.ctypes.data_as(P), ctypes.c_long(n), ctypes.c_long(total),
                       bitrate_kbit, no_gain, no_tonal, bfu, window_auto, window_mask,
                       out.ctypes.data_as(P), ctypes.c_long(out.size), sizes.ctypes.data_as(P),
                       ctypes.c_long(sizes.size), ctypes.byref(nb))
    assert k >= 0, "reference harness failed"
    return out[:nb.value].copy(), sizes[:k].copy()


def pad_units(payload, sizes, unit_bytes):
    """Container view of WriteFrame payloads: each resized to unit_bytes (src/aea.cpp:182, src/raw.cpp:41-43)."""
    o = np.zeros((len(sizes), unit_bytes), np.uint8)
    p = 0
    for i, s in enumerate(sizes):
        m = min(int(s), unit_bytes)
        o[i, :m] = payload[p:p + m]
        p += int(s)
    return o


def port_at1_encode(channels, pcm, window_auto=1, window_mask=0, bfu=0):
    lib = port_lib()
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    F = pcm.size // channels // 512
    out = np.zeros((F * channels, 212), np.uint8)
    sizes = np.zeros(F * channels, np.int32)
    lib.oat1_encode(channels, window_auto, window_mask, bfu, pcm.ctypes.data_as(P), ctypes.c_long(F),
                    out.ctypes.data_as(P), sizes.ctypes.data_as(P))
    return out, sizes


def ref_at1_stages(channels, pcm, window_auto=1, window_mask=0):
    lib = ref_lib()
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    F = pcm.size // channels // 512
    bands = np.zeros((F, channels, 512), np.float32)
    masks = np.zeros((F, channels), np.uint8)
    specs = np.zeros((F, channels, 512), np.float32)
    chloud = np.zeros((F, channels), np.float32)
    loud = np.zeros(F, np.float32)
    sfi = np.zeros((F, channels, 52), np.uint8)
    lib.ref_at1_stages(channels, pcm.ctypes.data_as(P), ctypes.c_long(F), window_auto, window_mask,
                       bands.ctypes.data_as(P), masks.ctypes.data_as(P), specs.ctypes.data_as(P),
                       chloud.ctypes.data_as(P), loud.ctypes.data_as(P), sfi.ctypes.data_as(P))
    return dict(bands=bands, masks=masks, specs=specs, chloud=chloud, loud=loud, sfi=sfi)


# ---------------------------------------------------------------------------------------------
# deterministic workload generator shared by tests and bench.py
def synth_streams(n_streams, n_frames, frame_samples, channels, seed=0xA7AC):
    """[S][F*frame_samples][C] float32: noise + per-stream sine, +20 dB bursts every 7th frame,
    one silent stream; int16-quantised."""
    rng = np.random.Generator(np.random.Philox(seed))
    n = n_frames * frame_samples
    t = np.arange(n, dtype=np.float64)
    out = np.empty((n_streams, n, channels), np.float32)
    for s in range(n_streams):
        f = 100.0 * (1 + s % 160)
        x = 0.025 * rng.uniform(-1, 1, (n, channels))
        for c in range(channels):
            x[:, c] += 0.03 * np.sin(2 * np.pi * f * t / 44100 + c)
        for fr in range(0, n_frames, 7):
            x[fr * frame_samples: fr * frame_samples + 64] *= 10.0
        if s % 61 == 60:
            x[:] = 0
        out[s] = quantise(x)
    return out


def engine_view(pcm, channels, step, total=None, buf_frames=4096, lookahead=0):
    """What the frame lambda actually receives when src/main.cpp:697-705 pumps `pcm` through
    TPCMEngine(4096) + TWav's reader: whole 4096-sample-frame reads; a short last read zeroes only
    (missing*channels) BYTES after the data (TPCMBuffer::Zero memsets len*NumChannels bytes,
    src/pcmengin.h:91-94) and leaves the rest of the buffer stale.  With a look-ahead codec
    (`lookahead` = number of calls that return LOOK_AHEAD) the engine "drains" at the end of the
    input: when the reader is exhausted but `total` is not reached it calls the lambda once more per
    pending look-ahead frame on the STALE buffer (src/pcmengin.h:157-182).  Returns interleaved PCM
    of every frame handed to the lambda, in order."""
    pcm = np.ascontiguousarray(pcm, dtype=np.float32).reshape(-1, channels)
    n = pcm.shape[0]
    total = n if total is None else total
    buf = np.zeros((buf_frames, channels), np.float32)
    pos, processed, frames = 0, 0, []
    pending, to_drain = lookahead, 0
    while total > processed:
        got = min(buf_frames, n - pos)
        drain = False
        if got == 0:
            if not to_drain:
                break
            drain = True
        else:
            buf[:got] = pcm[pos:pos + got]
            pos += got
            if got != buf_frames:
                flat = buf.reshape(-1).view(np.uint8)
                start = got * channels * 4
                flat[start:start + (buf_frames - got) * channels] = 0
        last_pos = 0
        for i in range(0, buf_frames - step + 1, step):
            frames.append(buf[i:i + step].copy())
            if pending:
                pending -= 1
                to_drain += 1
                continue
            last_pos += step
            if drain and to_drain:
                to_drain -= 1
                break
        processed += last_pos
    return np.concatenate(frames).reshape(-1)


def config1_sine():
    """BASELINE.json configs[0]: 1 s mono 44.1 kHz 1 kHz sine at 0.5 FS, int16-quantised."""
    n = np.arange(44100)
    return quantise(0.5 * np.sin(2 * np.pi * 1000 * n / 44100))


# ---------------------------------------------------------------------------------------------
# ATRAC3 reference taps (oracle/ref_harness_at3.cpp)
AT3_REC = np.dtype([
    ("loud_term", np.float32), ("gscale", np.float32, (4, 3)), ("n_points", np.int32, (4,)),
    ("points", np.int32, (4, 8, 2)), ("sfi", np.int32, (32,)), ("energy", np.float32, (32,)),
    ("n_tonal", np.int32), ("tonal", np.int32, (64, 4)), ("bands", np.float32, (4, 256)),
])


def ref_at3_stages(channels, pcm, bitrate_kbit=0, no_gain=0, no_tonal=0):
    """Drives the reference TAtrac3Encoder frame by frame.  Returns (recs[Fo][C], tracked[Fo], frames[Fo][FrameSz])."""
    lib = ref_lib()
    assert lib.ref_at3_rec_size() == AT3_REC.itemsize
    lib.ref_at3_stages.restype = ctypes.c_long
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    F = pcm.size // channels // 1024
    recs = np.zeros((F, channels), AT3_REC)
    tracked = np.zeros(F, np.float32)
    out = np.zeros(F * 1024, np.uint8)
    nb = ctypes.c_long()
    k = lib.ref_at3_stages(channels, pcm.ctypes.data_as(P), ctypes.c_long(F), bitrate_kbit, no_gain, no_tonal,
                           recs.ctypes.data_as(P), tracked.ctypes.data_as(P), out.ctypes.data_as(P),
                           ctypes.c_long(out.size), ctypes.byref(nb))
    assert k == F - 1 and nb.value > 0
    fs = nb.value // k
    return recs[:k], tracked[:k], out[:nb.value].reshape(k, fs)


def synth_rich(n_frames, frame_samples, channels, seed=1, kind="mix"):
    """One stream [n][C] of harder material than synth_streams: strong pure tones (tonal-component
    extraction), amplitude steps and clicks (gain control), silence gaps, near-full-scale passages."""
    rng = np.random.Generator(np.random.Philox(seed))
    n = n_frames * frame_samples
    t = np.arange(n, dtype=np.float64)
    x = np.zeros((n, channels))
    if kind == "tones":
        for c in range(channels):
            for k in range(4):
                f = rng.uniform(1500, 14000)
                x[:, c] += rng.uniform(0.05, 0.2) * np.sin(2 * np.pi * f * t / 44100 + rng.uniform(0, 6))
            x[:, c] += 1e-4 * rng.uniform(-1, 1, n)
    elif kind == "steps":
        env = np.ones(n)
        for k in range(0, n, frame_samples * 3):
            env[k + rng.integers(0, frame_samples):] *= rng.choice([0.05, 0.2, 4.0, 12.0])
            env = np.clip(env, 1e-3, 30)
        for c in range(channels):
            x[:, c] = 0.02 * env * (rng.uniform(-1, 1, n) + np.sin(2 * np.pi * 3000 * t / 44100 + c))
    else:
        for c in range(channels):
            x[:, c] = 0.02 * rng.uniform(-1, 1, n)
            for k in range(3):
                f = rng.uniform(200, 12000)
                x[:, c] += rng.uniform(0.01, 0.25) * np.sin(2 * np.pi * f * t / 44100 + rng.uniform(0, 6))
        for k in range(0, n_frames, 5):
            p = k * frame_samples + int(rng.integers(0, frame_samples))
            x[p:p + 200] += 0.5 * rng.uniform(-1, 1, (min(200, n - p), channels))
        for k in range(3, n_frames, 11):
            x[k * frame_samples:(k + 1) * frame_samples] = 0
        if channels == 2:
            x[n // 2:, 1] = 0.9 * x[n // 2:, 0] + 0.01 * rng.uniform(-1, 1, n - n // 2)   # correlated half: M/S budget shift
    return quantise(np.clip(x, -1, 1))


# ---------------------------------------------------------------------------------------------
# ATRAC3plus reference taps (oracle/ref_harness_at3p.cpp)
AT3P_GHA_REC = np.dtype([
    ("present", np.int32), ("num_tone_bands", np.int32), ("second_is_leader", np.int32),
    ("tone_sharing", np.int32, (16,)), ("n_sb", np.int32, (2,)), ("sb", np.int32, (2, 16, 4)),
    ("n_params", np.int32, (2,)), ("params", np.int32, (2, 64, 4)),
])


def ref_at3p_pqf(x):
    """at3plus_pqf_do_analyse over whole frames of one channel, fresh context -> [F][2048]."""
    lib = ref_lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    F = x.size // 2048
    out = np.zeros((F, 2048), np.float32)
    lib.ref_at3p_pqf(x.ctypes.data_as(P), ctypes.c_long(F), out.ctypes.data_as(P))
    return out


def ref_at3p_mdct(bands):
    """TAt3pMDCT::Do (sine windows, fresh history) over [F][2048] -> [F][2048]."""
    lib = ref_lib()
    bands = np.ascontiguousarray(bands, dtype=np.float32)
    F = bands.shape[0]
    out = np.zeros((F, 2048), np.float32)
    lib.ref_at3p_mdct(bands.ctypes.data_as(P), ctypes.c_long(F), out.ctypes.data_as(P))
    return out


def ref_at3p_stages(channels, pcm, gha_flags=-1):
    """Drives TAt3PEnc frame by frame; per output frame: the PQF-domain frames, the work buffer
    before / after the tone filter, the GHA result of that call, the spectra and the frame bytes."""
    lib = ref_lib()
    assert lib.ref_at3p_gha_rec_size() == AT3P_GHA_REC.itemsize
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    F = pcm.size // channels // 2048
    z = lambda: np.zeros((F, channels, 2048), np.float32)
    cur, nxt, win, wout, specs = z(), z(), z(), z(), z()
    gha = np.zeros(F, AT3P_GHA_REC)
    frames = np.zeros((F, 2048), np.uint8)
    lib.ref_at3p_stages.restype = ctypes.c_long
    k = lib.ref_at3p_stages(channels, pcm.ctypes.data_as(P), ctypes.c_long(F), gha_flags,
                            cur.ctypes.data_as(P), nxt.ctypes.data_as(P), win.ctypes.data_as(P),
                            wout.ctypes.data_as(P), gha.ctypes.data_as(P), specs.ctypes.data_as(P),
                            frames.ctypes.data_as(P))
    return dict(n=k, pqf_cur=cur[:k], pqf_next=nxt[:k], work_in=win[:k], work_out=wout[:k], gha=gha[:k],
                specs=specs[:k], frames=frames[:k])


def ref_at3p_pack(channels, specs, tones):
    """ScaleFrame + TAt3PBitStream::WriteFrame on caller-supplied spectra [U][C][2048] and tone records."""
    lib = ref_lib()
    specs = np.ascontiguousarray(specs, dtype=np.float32)
    tones = np.ascontiguousarray(tones)
    U = specs.shape[0]
    out = np.zeros((U, 2048), np.uint8)
    lib.ref_at3p_pack.restype = ctypes.c_long
    k = lib.ref_at3p_pack(channels, specs.ctypes.data_as(P), tones.ctypes.data_as(P), ctypes.c_long(U), out.ctypes.data_as(P))
    assert k == U
    return out
