"""Parity at benchmark scale: bench.py's own device-generated PCM goes through the reference encoder.

  * CPU: the verification plumbing (sampling, worker pool, unit comparison) on a tiny batch encoded by the kernel
    sources under the CPU emulator, incl. that a corrupted unit is reported;
  * GPU: `bench.py --verify` on BASELINE.json configs[2] (ATRAC3 LP2, 1024 x 977) and the per-GPU shard of configs[3]
    (LP4 joint stereo, 1024 x 1221): every frame of a 1-in-8 stream sample (>= 1.2*10^5 frames) equals the reference's.
"""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import atde_testlib as tl
import atracdenc_b200 as ab

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


@pytest.mark.parametrize("codec,kbit,step", [(3, 0, 1024), (3, 64, 1024), (1, 0, 512)])
def test_verify_plumbing_on_emulator(emu_lib, codec, kbit, step):
    import bench
    tl.require_ref("bench verification")
    S, F, C = 5, 4, 2
    pcm = tl.synth_streams(S, F, step, C, seed=4242)
    enc = ab.Encoder(codec, C, bitrate=kbit * 1024, lib=emu_lib)
    out = enc.encode(pcm, S)                                   # [S][Fo][units][unit_bytes]
    enc.close()
    out = out.reshape(S, -1, out.shape[-1])
    wl = dict(codec=codec, kbit=kbit)
    p = bench.verify_against_reference(wl, pcm, out)
    assert p["checked"] and p["mismatches"] == 0 and p["streams_checked"] == S
    assert p["frames_checked"] == S * (F - (1 if codec == 3 else 0))
    bad = out.copy()
    bad[3, 1, 7] ^= 0x10
    p = bench.verify_against_reference(wl, pcm, bad)
    assert p["mismatches"] == 1 and p["mismatching_streams"] == 1
    assert p["first_mismatch"] == {"sample_stream": 3, "unit": 1}


@pytest.mark.gpu
@pytest.mark.parametrize("workload", ["atrac3_lp2_stereo_1e6", "atrac3_lp4_stereo_1p25e6"])
def test_bench_batch_against_reference(workload):
    tl.require_ref("bench.py --verify")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--verify", "--workload", workload],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    p = line["parity"]
    assert p["checked"], p
    assert p["frames_checked"] >= 120000, p
    assert p["mismatches"] == 0, p
    assert line["host_and_device_outputs_equal"]
