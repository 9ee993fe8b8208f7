"""CPU-only: this repo's container writers (atracdenc_b200/host/atde_containers.cpp)
against the reference's own writers compiled unmodified into oracle/_ref — whole FILES must be byte-identical.
Covers every container main.cpp can select (AEA, raw, OMA for ATRAC3 and ATRAC3plus, RIFF/WAVE for both,
RealMedia) and the quirks that shape the bytes: AEA's swallowed first frame, dummy frame and 212-byte resize, the
title cut, raw's optional frame size, the OMA parameter word, the RIFF length fields rewritten from the delivered
frame count, RealMedia's three-frame packets, double-precision clock, scrambling and DATA size patch."""
import ctypes
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import atde_testlib as tl

ROOT = Path(__file__).resolve().parents[1]
AEA, RAW, OMA3, OMA3P, RIFF3, RIFF3P, RM = range(7)
ARGS = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]


@pytest.fixture(scope="module")
def writers():
    ref = tl.ref_lib()
    if ref is None or not hasattr(ref, "ref_container_write"):
        pytest.skip("oracle/_ref with the container writers not built (reference sources absent on this box)")
    out = ROOT / "tests" / "cpuemu" / "_build" / "libatde_containers_test.so"
    out.parent.mkdir(parents=True, exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(out),
                           str(ROOT / "tests" / "containers_driver.cpp"),
                           str(ROOT / "atracdenc_b200" / "host" / "atde_containers.cpp"),
                           "-I", str(ROOT / "atracdenc_b200" / "host")])
    ours = ctypes.CDLL(str(out))
    ref.ref_container_write.argtypes = ARGS
    ours.atde_container_write.argtypes = ARGS
    return ref.ref_container_write, ours.atde_container_write


def both(writers, tmp_path, kind, frames, title="test", channels=2, num_frames=None, frame_size=0, js=0, expect_rc=0):
    """frames: list of byte strings (the WriteFrame payloads).  Returns the common file image."""
    sizes = np.array([len(f) for f in frames], np.int32)
    blob = np.frombuffer(b"".join(frames) + b"\0", np.uint8).copy()
    n_est = len(frames) if num_frames is None else num_frames
    images = []
    for tag, fn in zip(("ref", "ours"), writers):
        path = tmp_path / f"{tag}_{kind}.bin"
        if tag == "ref" and kind == RM:
            # the reference scrambles into a FUNCTION-STATIC buffer sized by the first frame the process ever
            # writes (src/rm.cpp: `static std::vector<char> tmp(data.size())`); the CLI writes one file per
            # process, so every reference RealMedia file is produced in a process of its own
            blob_path = tmp_path / "rm_payload.bin"
            blob_path.write_bytes(blob.tobytes())
            code = (
                "import ctypes, sys, numpy as np\n"
                f"lib = ctypes.CDLL({str(tl.REF_SO)!r})\n"
                f"blob = np.fromfile({str(blob_path)!r}, np.uint8)\n"
                f"sizes = np.array({sizes.tolist()!r}, np.int32)\n"
                "lib.ref_container_write.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, "
                "ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]\n"
                f"sys.exit(lib.ref_container_write({kind}, {str(path).encode()!r}, {title.encode()!r}, {channels}, {n_est}, "
                f"{frame_size}, {js}, blob.ctypes.data, sizes.ctypes.data, {len(frames)}))\n")
            rc = subprocess.run([sys.executable, "-c", code]).returncode
        else:
            rc = fn(kind, str(path).encode(), title.encode(), channels, n_est, frame_size, js,
                    blob.ctypes.data, sizes.ctypes.data, len(frames))
        assert rc == expect_rc, (tag, rc)
        images.append(path.read_bytes())
    assert images[0] == images[1], f"kind {kind}: files differ (sizes {len(images[0])} / {len(images[1])})"
    return images[0]


def payloads(rng, n, size, jitter=0):
    return [rng.integers(0, 256, size + (int(rng.integers(-jitter, jitter + 1)) if jitter else 0), dtype=np.uint8).tobytes()
            for _ in range(n)]


def test_aea(writers, tmp_path):
    rng = np.random.default_rng(1)
    img = both(writers, tmp_path, AEA, payloads(rng, 9, 212), channels=2)
    assert len(img) == 2048 + 212 + 8 * 212                      # first payload swallowed, dummy frame in front
    both(writers, tmp_path, AEA, payloads(rng, 7, 212, jitter=9), channels=1, title="a title longer than 16 chars", num_frames=1234567)
    both(writers, tmp_path, AEA, [], channels=2, title="")
    both(writers, tmp_path, AEA, payloads(rng, 1, 100), channels=2, title="exactly15chars!")


def test_raw(writers, tmp_path):
    rng = np.random.default_rng(2)
    both(writers, tmp_path, RAW, payloads(rng, 6, 212, jitter=20), frame_size=212)      # ATRAC1: SoundUnitSize
    img = both(writers, tmp_path, RAW, payloads(rng, 6, 384, jitter=20), frame_size=0)  # ATRAC3 / ATRAC3plus: verbatim
    both(writers, tmp_path, RAW, [], frame_size=0)


@pytest.mark.parametrize("frame_size,js", [(384, 0), (192, 1), (304, 0), (512, 0), (8192, 0)])
def test_oma_atrac3(writers, tmp_path, frame_size, js):
    rng = np.random.default_rng(3)
    img = both(writers, tmp_path, OMA3, payloads(rng, 5, frame_size), frame_size=frame_size, js=js)
    assert len(img) == 96 + 5 * frame_size


@pytest.mark.parametrize("channels", [1, 2])
def test_oma_atrac3plus(writers, tmp_path, channels):
    rng = np.random.default_rng(4)
    both(writers, tmp_path, OMA3P, payloads(rng, 4, 2048), channels=channels, frame_size=2048)
    both(writers, tmp_path, OMA3P, [], channels=channels, frame_size=2048)


@pytest.mark.parametrize("frame_size,js", [(384, 0), (192, 1), (424, 0)])
def test_riff_atrac3(writers, tmp_path, frame_size, js):
    rng = np.random.default_rng(5)
    # the estimate handed to the constructor is one frame short of what the engine delivers (look-ahead tail)
    img = both(writers, tmp_path, RIFF3, payloads(rng, 11, frame_size), num_frames=10, frame_size=frame_size, js=js)
    assert len(img) == 76 + 11 * frame_size
    assert int.from_bytes(img[72:76], "little") == 11 * frame_size
    both(writers, tmp_path, RIFF3, [], num_frames=10, frame_size=frame_size, js=js)      # nothing delivered: estimate stays
    both(writers, tmp_path, RIFF3, payloads(rng, 3, frame_size, jitter=5), num_frames=3, frame_size=frame_size, js=js)
    both(writers, tmp_path, RIFF3, [], num_frames=0xFFFFFFFF // frame_size + 1, frame_size=frame_size, expect_rc=1)  # too big


@pytest.mark.parametrize("channels", [1, 2])
def test_riff_atrac3plus(writers, tmp_path, channels):
    rng = np.random.default_rng(6)
    img = both(writers, tmp_path, RIFF3P, payloads(rng, 6, 2048), channels=channels, num_frames=5, frame_size=2048)
    assert len(img) == 80 + 6 * 2048
    both(writers, tmp_path, RIFF3P, [], channels=channels, num_frames=7, frame_size=2048)
    # a payload of the wrong size is refused by both (the frames before it are on disk)
    both(writers, tmp_path, RIFF3P, payloads(rng, 2, 2048) + payloads(rng, 1, 2047), channels=channels, frame_size=2048, expect_rc=1)


@pytest.mark.parametrize("frame_size,js,n", [(384, 0, 9), (192, 1, 10), (304, 0, 11), (384, 0, 0), (512, 0, 1)])
def test_realmedia(writers, tmp_path, frame_size, js, n):
    rng = np.random.default_rng(7)
    img = both(writers, tmp_path, RM, payloads(rng, n, frame_size), num_frames=max(n, 1) + 2, frame_size=frame_size, js=js)
    assert img[:4] == b".RMF"
    assert len(img) == 18 + 50 + 168 + 18 + n * frame_size + 12 * ((n + 2) // 3)


def test_realmedia_long_clock(writers, tmp_path):
    """the packet timestamps come from a double accumulated in steps of three frame durations and truncated"""
    rng = np.random.default_rng(8)
    both(writers, tmp_path, RM, payloads(rng, 400, 192), num_frames=400, frame_size=192, js=1)
