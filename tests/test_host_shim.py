"""The C++ host shim (atracdenc_b200/host) mirrors TAtrac1Encoder's interface over the C ABI.
CPU: built against the emulated kernel library; GPU: against libatde_b200.so.  The driver mimics
src/main.cpp's PCM loop; the captured WriteFrame payloads (bytes AND lengths, in order) must equal
the reference's."""
import subprocess
import sys
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
                           str(ROOT / "atracdenc_b200" / "host" / "atde_encoders.cpp"),
                           str(ROOT / "atracdenc_b200" / "host" / "atde_batcher.cpp"),
                           str(ROOT / "atracdenc_b200" / "host" / "atde_gain_trace.cpp"),
                           str(ROOT / "atracdenc_b200" / "host" / "atde_containers.cpp"), str(so),
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


# ---- whole files: shim + this repo's container writers vs reference encoder + reference writers --------------
CONTAINER_KIND = {(1, "aea"): 0, (1, "raw"): 1, (3, "oma"): 2, (3, "riff"): 4, (3, "rm"): 6, (3, "raw"): 1,
                  (4, "oma"): 3, (4, "riff"): 5, (4, "raw"): 1}


def check_files(exe, tmp_path, codec, containers, seconds=0.3, kbit=0):
    """What `atracdenc -e ... -i in.wav -o out.<ext>` writes: the reference encoder's WriteFrame calls fed to the
    reference's container writer (main.cpp's constructor arguments) against the shim feeding this repo's writer."""
    import ctypes
    ref = tl.ref_lib()
    if ref is None or not hasattr(ref, "ref_container_write"):
        pytest.skip("oracle/_ref with the container writers not built")
    ref.ref_container_write.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_uint32,
                                        ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    step = {1: 512, 3: 1024, 4: 2048}[codec]
    ch = 2
    n = int(44100 * seconds)
    pcm = tl.synth_rich((n + step - 1) // step, step, ch, seed=81 + codec)[:n]
    payload, sizes = tl.ref_encode(codec, ch, pcm.reshape(-1), total=n, bitrate_kbit=kbit)
    payload = np.ascontiguousarray(payload, np.uint8)
    sizes = np.ascontiguousarray(sizes, np.int32)
    frame_size, js = {1: (212, 0), 4: (2048, 0)}.get(codec, ((192, 1) if kbit == 64 else (384, 0)))
    num_frames = ch * n // 512 if codec == 1 else n // step
    src, batch = tmp_path / "in.f32", 5
    np.ascontiguousarray(pcm, np.float32).tofile(src)
    for cont in containers:
        kind = CONTAINER_KIND[(codec, cont)]
        fs = 212 if (codec == 1 and cont == "raw") else (0 if cont == "raw" else frame_size)
        want_path, got_path = tmp_path / f"ref_{codec}_{cont}.bin", tmp_path / f"got_{codec}_{cont}.bin"
        if cont == "rm":
            # the reference's RealMedia writer keeps a function-static scramble buffer sized by the first frame of
            # the PROCESS (src/rm.cpp); like the CLI, give every file a process of its own
            blob_path, sizes_path = tmp_path / "rm_payload.bin", tmp_path / "rm_sizes.bin"
            payload.tofile(blob_path); sizes.tofile(sizes_path)
            code = ("import ctypes, sys, numpy as np\n"
                    f"lib = ctypes.CDLL({str(tl.REF_SO)!r})\n"
                    f"blob = np.fromfile({str(blob_path)!r}, np.uint8); sizes = np.fromfile({str(sizes_path)!r}, np.int32)\n"
                    "lib.ref_container_write.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, "
                    "ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]\n"
                    f"sys.exit(lib.ref_container_write({kind}, {str(want_path).encode()!r}, b'test', {ch}, {num_frames}, {fs}, {js}, "
                    "blob.ctypes.data, sizes.ctypes.data, len(sizes)))\n")
            rc = subprocess.run([sys.executable, "-c", code]).returncode
        else:
            rc = ref.ref_container_write(kind, str(want_path).encode(), b"test", 2 if (codec == 3 and cont == "riff") else ch,
                                         num_frames, fs, js, payload.ctypes.data, sizes.ctypes.data, len(sizes))
        assert rc == 0
        subprocess.check_call([str(exe), str(src), str(ch), str(n), str(got_path), str(batch), str(codec), str(kbit), cont])
        assert got_path.read_bytes() == want_path.read_bytes(), (codec, cont)


def test_whole_files_cpu_emulated(tmp_path):
    exe = build_driver(tl.build_emu(), "emu")
    check_files(exe, tmp_path, 1, ["aea", "raw"], seconds=0.2)
    check_files(exe, tmp_path, 3, ["oma", "riff", "rm", "raw"], seconds=0.2)
    check_files(exe, tmp_path, 3, ["oma", "riff", "rm"], seconds=0.2, kbit=64)
    check_files(exe, tmp_path, 4, ["oma", "riff", "raw"], seconds=0.2)


@pytest.mark.gpu
def test_whole_files_gpu(tmp_path, gpu_lib):
    exe = build_driver(ROOT / "atracdenc_b200" / "libatde_b200.so", "gpu")
    check_files(exe, tmp_path, 1, ["aea", "raw"], seconds=2.0)
    check_files(exe, tmp_path, 3, ["oma", "riff", "rm", "raw"], seconds=2.0)
    check_files(exe, tmp_path, 3, ["oma", "riff", "rm"], seconds=1.0, kbit=64)
    check_files(exe, tmp_path, 4, ["oma", "riff", "raw"], seconds=1.0)
