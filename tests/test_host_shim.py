"""The C++ host shim (atracdenc_b200/host) mirrors TAtrac1Encoder's interface over the C ABI.
CPU: built against the emulated kernel library; GPU: against libatde_b200.so.  The driver mimics
src/main.cpp's PCM loop; the captured WriteFrame payloads (bytes AND lengths, in order) must equal
the reference's."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import atde_testlib as tl

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


def build_driver(so: Path, tag: str) -> Path:
    out = ROOT / "tests" / "cpuemu" / "_build" / f"host_shim_driver_{tag}"
    out.parent.mkdir(parents=True, exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(out), str(ROOT / "tests" / "host_shim_driver.cpp"),
                           str(ROOT / "atracdenc_b200" / "host" / "atde_encoders.cpp"), str(so),
                           f"-Wl,-rpath,{so.parent}", "-pthread"])
    return out


def run_driver(exe, pcm, channels, total, batch, tmp_path, codec=1, kbit=0):
    src, dst = tmp_path / "in.f32", tmp_path / "out.bin"
    np.ascontiguousarray(pcm, np.float32).tofile(src)
    subprocess.check_call([str(exe), str(src), str(channels), str(total), str(dst), str(batch), str(codec), str(kbit)])
    raw = dst.read_bytes()
    sizes, payload, p = [], bytearray(), 0
    while p < len(raw):
        n = int(np.frombuffer(raw[p:p + 4], np.int32)[0]); p += 4
        sizes.append(n); payload += raw[p:p + n]; p += n
    return np.frombuffer(bytes(payload), np.uint8), np.array(sizes, np.int32)


def check(exe, tmp_path):
    g = np.load(GOLDEN / "at1_config1_sine_mono.npz")
    for batch in (7, 4096):
        payload, sizes = run_driver(exe, tl.config1_sine(), 1, 44100, batch, tmp_path)
        assert np.array_equal(sizes, g["sizes"])
        assert np.array_equal(tl.pad_units(payload, sizes, 212), g["units"])


def check_at3p(exe, tmp_path, seconds=0.6):
    """TAt3PEnc mirror under main.cpp's PCM pump (incl. the look-ahead drain) against the reference."""
    if tl.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    n = int(44100 * seconds)
    pcm = tl.synth_rich((n + 2047) // 2048, 2048, 2, seed=71)[:n]
    want, want_sizes = tl.ref_encode(4, 2, pcm.reshape(-1), total=n)
    for batch in (3, 4096):
        payload, sizes = run_driver(exe, pcm, 2, n, batch, tmp_path, codec=4)
        assert np.array_equal(sizes, want_sizes)
        assert np.array_equal(payload, want)


def check_at3(exe, tmp_path, seconds=0.5):
    """TAtrac3Encoder mirror under main.cpp's PCM pump (incl. the look-ahead drain at the end of the
    input) against the reference encoder under the same pump."""
    if tl.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    n = int(44100 * seconds)
    pcm = tl.synth_rich((n + 1023) // 1024, 1024, 2, seed=61)[:n]
    for kbit in (0, 64):
        want, want_sizes = tl.ref_encode(3, 2, pcm.reshape(-1), total=n, bitrate_kbit=kbit)
        for batch in (5, 4096):
            payload, sizes = run_driver(exe, pcm, 2, n, batch, tmp_path, codec=3, kbit=kbit)
            assert np.array_equal(sizes, want_sizes)
            assert np.array_equal(payload, want)


def test_host_shim_cpu_emulated(tmp_path):
    check(build_driver(tl.build_emu(), "emu"), tmp_path)


def test_host_shim_at3_cpu_emulated(tmp_path):
    check_at3(build_driver(tl.build_emu(), "emu"), tmp_path, seconds=0.3)


@pytest.mark.gpu
def test_host_shim_at3_gpu(tmp_path, gpu_lib):
    check_at3(build_driver(ROOT / "atracdenc_b200" / "libatde_b200.so", "gpu"), tmp_path, seconds=3.0)


@pytest.mark.gpu
def test_host_shim_gpu(tmp_path, gpu_lib):
    check(build_driver(ROOT / "atracdenc_b200" / "libatde_b200.so", "gpu"), tmp_path)


def test_host_shim_at3p_cpu_emulated(tmp_path):
    check_at3p(build_driver(tl.build_emu(), "emu"), tmp_path, seconds=0.4)


@pytest.mark.gpu
def test_host_shim_at3p_gpu(tmp_path, gpu_lib):
    check_at3p(build_driver(ROOT / "atracdenc_b200" / "libatde_b200.so", "gpu"), tmp_path, seconds=2.0)
