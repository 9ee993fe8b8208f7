"""
atracdenc_b200 — ctypes binding of libatde_b200.so (include/atde_b200.h).

This module is harness glue for tests/ and bench.py; the product is the shared library and the
C++ host shim under atracdenc_b200/host/.  It never computes anything itself and it never falls
back to a CPU implementation: if the CUDA library is missing or no GPU is visible, it raises.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libatde_b200.so"

CODEC_ATRAC1 = 1
CODEC_ATRAC3 = 3
CODEC_ATRAC3PLUS = 4

TAP_SPECS, TAP_MASKS, TAP_CHLOUD, TAP_LOUDNESS, TAP_SFI, TAP_WORDLEN = 1, 2, 3, 4, 5, 6
TAP_BANDS, TAP_CURVES, TAP_GSCALE, TAP_ENERGY, TAP_TONAL, TAP_GAIN = 7, 8, 9, 10, 11, 12
TAP_TRACE_GAIN, TAP_TRACE_STAT = 13, 14      # with Encoder.set_gain_trace(True): the envelope analysis over all four bands

EXPORTS = [
    "atde_default_settings", "atde_create", "atde_destroy", "atde_frame_samples",
    "atde_units_per_frame", "atde_unit_bytes", "atde_lookahead_frames", "atde_output_frames", "atde_encode_batch", "atde_encode_batch_i16",
    "atde_encode_batch_device", "atde_sync", "atde_reset", "atde_cuda_stream",
    "atde_launch_count", "atde_set_profiling", "atde_set_gain_trace", "atde_kernel_times", "atde_debug_tap", "atde_debug_math", "atde_last_error", "atde_version",
    "atde_create_group", "atde_destroy_group", "atde_group_size", "atde_group_encode_batch", "atde_group_encode_batch_i16",
    "atde_group_output_frames", "atde_group_reset",
    "atde_decoder_create", "atde_decoder_destroy", "atde_decode_batch", "atde_decoder_reset",
]


class Settings(ctypes.Structure):
    _fields_ = [
        ("codec", ctypes.c_int32), ("channels", ctypes.c_int32),
        ("bfu_idx_const", ctypes.c_uint32), ("window_mode", ctypes.c_int32),
        ("window_mask", ctypes.c_uint32), ("bitrate", ctypes.c_uint32),
        ("no_gain_control", ctypes.c_int32), ("no_tonal", ctypes.c_int32),
        ("device", ctypes.c_int32), ("gha_flags", ctypes.c_uint32), ("reserved", ctypes.c_int32 * 6),
    ]


class AtdeError(RuntimeError):
    pass


def load_library(path: os.PathLike | str | None = None) -> ctypes.CDLL:
    """Loads the C-ABI library and declares the prototypes of include/atde_b200.h."""
    p = Path(path) if path else Path(os.environ.get("ATDE_LIB", LIB_PATH))   # ATDE_LIB: A/B a differently built library
    if not p.exists():
        raise AtdeError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    lib = ctypes.CDLL(str(p))
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.atde_default_settings.argtypes = [ctypes.POINTER(Settings), i32, i32]
    lib.atde_default_settings.restype = None
    lib.atde_create.argtypes = [ctypes.POINTER(Settings), ctypes.POINTER(vp)]
    lib.atde_destroy.argtypes = [vp]
    lib.atde_destroy.restype = None
    for f in ("atde_frame_samples", "atde_units_per_frame", "atde_unit_bytes", "atde_lookahead_frames",
              "atde_sync", "atde_reset"):
        getattr(lib, f).argtypes = [vp]
        getattr(lib, f).restype = ctypes.c_int
    lib.atde_output_frames.argtypes = [vp, i64]
    lib.atde_output_frames.restype = i64
    lib.atde_encode_batch.argtypes = [vp, vp, i32, i64, vp, vp]
    lib.atde_encode_batch_i16.argtypes = [vp, vp, i32, i64, vp, vp]
    lib.atde_encode_batch_device.argtypes = [vp, vp, i32, i64, vp, vp]
    lib.atde_cuda_stream.argtypes = [vp]
    lib.atde_cuda_stream.restype = vp
    lib.atde_launch_count.argtypes = [vp]
    lib.atde_launch_count.restype = i64
    lib.atde_set_profiling.argtypes = [vp, i32]
    lib.atde_set_gain_trace.argtypes = [vp, i32]
    lib.atde_kernel_times.argtypes = [vp, vp, vp, i32]
    lib.atde_debug_tap.argtypes = [vp, i32, vp, ctypes.c_size_t]
    lib.atde_debug_tap.restype = i64
    lib.atde_debug_math.argtypes = [i32, i32, vp, vp, i64]
    lib.atde_create_group.argtypes = [ctypes.POINTER(Settings), ctypes.POINTER(i32), i32, ctypes.POINTER(vp)]
    lib.atde_destroy_group.argtypes = [vp]
    lib.atde_destroy_group.restype = None
    lib.atde_group_size.argtypes = [vp]
    lib.atde_group_encode_batch.argtypes = [vp, vp, i32, i64, vp, vp]
    lib.atde_group_encode_batch_i16.argtypes = [vp, vp, i32, i64, vp, vp]
    lib.atde_group_output_frames.argtypes = [vp, i64]
    lib.atde_group_output_frames.restype = i64
    lib.atde_group_reset.argtypes = [vp]
    lib.atde_decoder_create.argtypes = [i32, i32, ctypes.POINTER(vp)]
    lib.atde_decoder_destroy.argtypes = [vp]
    lib.atde_decoder_destroy.restype = None
    lib.atde_decode_batch.argtypes = [vp, vp, i32, i64, vp]
    lib.atde_decoder_reset.argtypes = [vp]
    lib.atde_last_error.restype = ctypes.c_char_p
    lib.atde_version.restype = ctypes.c_char_p
    return lib


class Encoder:
    """One atde_encoder handle == one batch of reference encoder instances (streams)."""

    def __init__(self, codec: int, channels: int, *, bfu_idx_const: int = 0, window_mode: int = 1,
                 window_mask: int = 0, bitrate: int = 0, no_gain_control: bool = False,
                 no_tonal: bool = False, device: int = 0, lib: ctypes.CDLL | None = None, gha_flags: int | None = None):
        self.lib = lib or load_library()
        s = Settings()
        self.lib.atde_default_settings(ctypes.byref(s), codec, channels)
        s.bfu_idx_const, s.window_mode, s.window_mask = bfu_idx_const, window_mode, window_mask
        s.bitrate, s.no_gain_control, s.no_tonal, s.device = bitrate, int(no_gain_control), int(no_tonal), device
        if gha_flags is not None:
            s.gha_flags = gha_flags
        h = ctypes.c_void_p()
        self._check(self.lib.atde_create(ctypes.byref(s), ctypes.byref(h)))
        self.h = h
        self.channels = channels
        self.frame_samples = self.lib.atde_frame_samples(h)
        self.units_per_frame = self.lib.atde_units_per_frame(h)
        self.unit_bytes = self.lib.atde_unit_bytes(h)
        self.lookahead = self.lib.atde_lookahead_frames(h)

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise AtdeError(f"atde error {rc}: {self.lib.atde_last_error().decode()}")
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.atde_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self):
        self._check(self.lib.atde_reset(self.h))

    def arm_taps(self):
        dummy = ctypes.c_int()
        self._check(self.lib.atde_debug_tap(self.h, 0, ctypes.byref(dummy), 0))

    def encode(self, pcm: np.ndarray, n_streams: int, want_sizes: bool = False):
        """pcm: float32 [S][F*frame_samples][C] (any shape with that memory order).
        Returns uint8 [S][Fo][units][unit_bytes] (and int32 sizes [S][Fo][units]), Fo = output_frames(F)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        per_stream = pcm.size // n_streams
        assert per_stream * n_streams == pcm.size
        F = per_stream // (self.frame_samples * self.channels)
        assert F * self.frame_samples * self.channels == per_stream, "whole frames only"
        Fo = self.output_frames(F)
        out = np.empty((n_streams, Fo, self.units_per_frame, self.unit_bytes), dtype=np.uint8)
        sizes = np.empty((n_streams, Fo, self.units_per_frame), dtype=np.int32) if want_sizes else None
        self._check(self.lib.atde_encode_batch(
            self.h, pcm.ctypes.data, n_streams, F, out.ctypes.data,
            sizes.ctypes.data if want_sizes else None))
        return (out, sizes) if want_sizes else out

    def encode_i16(self, pcm: np.ndarray, n_streams: int, want_sizes: bool = False):
        """Like encode(), from int16 PCM [S][F*frame_samples][C] (converted on the device, value / 32768)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        per_stream = pcm.size // n_streams
        assert per_stream * n_streams == pcm.size
        F = per_stream // (self.frame_samples * self.channels)
        assert F * self.frame_samples * self.channels == per_stream, "whole frames only"
        Fo = self.output_frames(F)
        out = np.empty((n_streams, Fo, self.units_per_frame, self.unit_bytes), dtype=np.uint8)
        sizes = np.empty((n_streams, Fo, self.units_per_frame), dtype=np.int32) if want_sizes else None
        self._check(self.lib.atde_encode_batch_i16(
            self.h, pcm.ctypes.data, n_streams, F, out.ctypes.data,
            sizes.ctypes.data if want_sizes else None))
        return (out, sizes) if want_sizes else out

    def encode_ptr_i16(self, pcm_ptr: int, n_streams: int, n_frames: int, out_ptr: int, sizes_ptr: int = 0):
        """Host-pointer variant of encode_i16 (pinned buffers owned by the caller)."""
        self._check(self.lib.atde_encode_batch_i16(self.h, pcm_ptr, n_streams, n_frames, out_ptr, sizes_ptr or None))

    def output_frames(self, n_frames: int) -> int:
        """Output frames per stream the next batch of n_frames will produce (ATRAC3's first batch: n-1)."""
        return self._check(self.lib.atde_output_frames(self.h, n_frames))

    def encode_ptr(self, pcm_ptr: int, n_streams: int, n_frames: int, out_ptr: int, sizes_ptr: int = 0):
        """Host-pointer variant (pinned buffers owned by the caller)."""
        self._check(self.lib.atde_encode_batch(self.h, pcm_ptr, n_streams, n_frames, out_ptr, sizes_ptr or None))

    def encode_device(self, d_pcm: int, n_streams: int, n_frames: int, d_out: int, d_sizes: int = 0):
        self._check(self.lib.atde_encode_batch_device(self.h, d_pcm, n_streams, n_frames, d_out, d_sizes or None))

    def sync(self):
        self._check(self.lib.atde_sync(self.h))

    @property
    def cuda_stream(self) -> int:
        return self.lib.atde_cuda_stream(self.h) or 0

    @property
    def launch_count(self) -> int:
        return self.lib.atde_launch_count(self.h)

    def set_profiling(self, on: bool):
        self._check(self.lib.atde_set_profiling(self.h, int(on)))

    def set_gain_trace(self, on: bool):
        """ATRAC3: keep the data of the reference's `--yaml-log` gain-control trace (atde_set_gain_trace)."""
        self._check(self.lib.atde_set_gain_trace(self.h, int(on)))

    def kernel_times(self, n_kinds: int = 3):
        """(ms_sum[kind], count[kind]) since the last query; kinds: 0 analysis, 1 loudness, 2 pack."""
        ms = (ctypes.c_double * n_kinds)()
        cnt = (ctypes.c_int64 * n_kinds)()
        self._check(self.lib.atde_kernel_times(self.h, ms, cnt, n_kinds))
        return list(ms), list(cnt)

    def tap(self, what: int, shape, dtype):
        buf = np.empty(shape, dtype=dtype)
        self._check(self.lib.atde_debug_tap(self.h, what, buf.ctypes.data, buf.nbytes))
        return buf


class EncoderGroup:
    """atde_create_group: one member per device, a batch's streams sharded contiguously over the members, members run
    concurrently (one host thread per device).  Host buffers only."""

    def __init__(self, codec: int, channels: int, devices, *, bitrate: int = 0, lib: ctypes.CDLL | None = None, **kw):
        self.lib = lib or load_library()
        s = Settings()
        self.lib.atde_default_settings(ctypes.byref(s), codec, channels)
        s.bitrate = bitrate
        for k, v in kw.items():
            setattr(s, k, int(v))
        devs = (ctypes.c_int32 * len(devices))(*devices)
        h = ctypes.c_void_p()
        rc = self.lib.atde_create_group(ctypes.byref(s), devs, len(devices), ctypes.byref(h))
        if rc < 0:
            raise AtdeError(f"atde error {rc}: {self.lib.atde_last_error().decode()}")
        self.h, self.channels = h, channels
        probe = Encoder(codec, channels, bitrate=bitrate, device=devices[0], lib=self.lib)
        self.frame_samples, self.units_per_frame, self.unit_bytes = probe.frame_samples, probe.units_per_frame, probe.unit_bytes
        probe.close()

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise AtdeError(f"atde error {rc}: {self.lib.atde_last_error().decode()}")
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.atde_destroy_group(self.h)
            self.h = None

    __del__ = close

    def size(self) -> int:
        return self._check(self.lib.atde_group_size(self.h))

    def reset(self):
        self._check(self.lib.atde_group_reset(self.h))

    def output_frames(self, n_frames: int) -> int:
        return self._check(self.lib.atde_group_output_frames(self.h, n_frames))

    def encode(self, pcm: np.ndarray, n_streams: int, want_sizes: bool = False):
        i16 = pcm.dtype == np.int16
        pcm = np.ascontiguousarray(pcm, dtype=np.int16 if i16 else np.float32)
        per_stream = pcm.size // n_streams
        F = per_stream // (self.frame_samples * self.channels)
        assert F * self.frame_samples * self.channels * n_streams == pcm.size, "whole frames only"
        Fo = self.output_frames(F)
        out = np.empty((n_streams, Fo, self.units_per_frame, self.unit_bytes), dtype=np.uint8)
        sizes = np.empty((n_streams, Fo, self.units_per_frame), dtype=np.int32) if want_sizes else None
        fn = self.lib.atde_group_encode_batch_i16 if i16 else self.lib.atde_group_encode_batch
        self._check(fn(self.h, pcm.ctypes.data, n_streams, F, out.ctypes.data, sizes.ctypes.data if want_sizes else None))
        return (out, sizes) if want_sizes else out

    def encode_ptr(self, pcm_ptr: int, n_streams: int, n_frames: int, out_ptr: int, i16: bool = False):
        fn = self.lib.atde_group_encode_batch_i16 if i16 else self.lib.atde_group_encode_batch
        self._check(fn(self.h, pcm_ptr, n_streams, n_frames, out_ptr, None))


class Decoder:
    """ATRAC1 decoder handle (atde_decoder_create): sound units [S][F][C][212] -> PCM [S][F*512][C]."""

    def __init__(self, channels: int, device: int = 0, lib: ctypes.CDLL | None = None):
        self.lib = lib or load_library()
        h = ctypes.c_void_p()
        rc = self.lib.atde_decoder_create(channels, device, ctypes.byref(h))
        if rc < 0:
            raise AtdeError(f"atde error {rc}: {self.lib.atde_last_error().decode()}")
        self.h, self.channels = h, channels

    def close(self):
        if getattr(self, "h", None):
            self.lib.atde_decoder_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self):
        self.lib.atde_decoder_reset(self.h)

    def decode(self, units: np.ndarray, n_streams: int) -> np.ndarray:
        units = np.ascontiguousarray(units, dtype=np.uint8)
        F = units.size // (n_streams * self.channels * 212)
        assert F * n_streams * self.channels * 212 == units.size, "whole sound units only"
        pcm = np.empty((n_streams, F * 512, self.channels), dtype=np.float32)
        rc = self.lib.atde_decode_batch(self.h, units.ctypes.data, n_streams, F, pcm.ctypes.data)
        if rc < 0:
            raise AtdeError(f"atde error {rc}: {self.lib.atde_last_error().decode()}")
        return pcm
