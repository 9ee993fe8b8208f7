"""GPU parity suite, ATRAC3: libatde_b200.so through the C ABI on a real B200 vs the reference
encoder (oracle/_ref) and the committed golden fixtures."""
import numpy as np
import pytest

import atde_testlib as tl
import atracdenc_b200 as ab
import parity_cases as pc

pytestmark = pytest.mark.gpu


def test_golden_lp2(gpu_lib):
    pc.check_at3_golden(gpu_lib, "at3_lp2_stereo.npz")


def test_golden_lp4(gpu_lib):
    pc.check_at3_golden(gpu_lib, "at3_lp4_js_stereo.npz")


@pytest.mark.parametrize("kbit,C", [(0, 2), (64, 2), (0, 1), (128, 2), (94, 2), (64, 1), (90, 1)])
def test_vs_oracle(gpu_lib, kbit, C):
    """LP2, LP4 (joint stereo), mono LP2, and two more containers (132300 via 128 kbit request;
    104738 non-JS via 94*1024=96256)."""
    n = pc.check_at3_vs_oracle(gpu_lib, S=8, F=48, C=C, kbit=kbit, seed=1000 + kbit)
    assert n == 8


def test_vs_oracle_flags(gpu_lib):
    pc.check_at3_vs_oracle(gpu_lib, S=4, F=30, C=2, kbit=0, seed=400, no_gain=1)
    pc.check_at3_vs_oracle(gpu_lib, S=4, F=30, C=2, kbit=0, seed=410, no_tonal=1)
    pc.check_at3_vs_oracle(gpu_lib, S=4, F=30, C=2, kbit=64, seed=420, no_gain=1, no_tonal=1)


def test_stage_taps(gpu_lib):
    pc.check_at3_stage_taps(gpu_lib, C=2, F=40, kbit=0)
    pc.check_at3_stage_taps(gpu_lib, C=2, F=40, kbit=64, kind="steps")
    pc.check_at3_stage_taps(gpu_lib, C=1, F=30, kbit=0, kind="tones")


def test_main_loop_view(gpu_lib):
    pc.check_at3_main_loop(gpu_lib, C=2, kbit=0, seconds=2.0)
    pc.check_at3_main_loop(gpu_lib, C=2, kbit=64, seconds=1.0)


def test_batch_split_invariance(gpu_lib):
    pc.check_at3_batch_split_invariance(gpu_lib, S=4, F=37, C=2, kbit=0, cuts=(9, 14))
    pc.check_at3_batch_split_invariance(gpu_lib, S=3, F=20, C=2, kbit=64, cuts=(1, 1))


def test_stream_independence(gpu_lib):
    pc.check_at3_stream_independence(gpu_lib, F=17)
    pc.check_at3_stream_independence(gpu_lib, F=17, kbit=64)


def test_edge_inputs(gpu_lib):
    pc.check_at3_edge_inputs(gpu_lib)
    pc.check_at3_edge_inputs(gpu_lib, kbit=64)


def test_errors(gpu_lib):
    pc.check_at3_errors(gpu_lib)


def test_long_streams_against_oracle(gpu_lib):
    """64 streams x 200 frames per mode (25 k stereo frames): every stream against the reference."""
    for kbit in (0, 64):
        n = pc.check_at3_vs_oracle(gpu_lib, S=64, F=200, C=2, kbit=kbit, seed=5000 + kbit)
        assert n == 64


@pytest.mark.parametrize("kbit", [0, 64])
def test_full_size_properties(gpu_lib, kbit):
    """BASELINE.json configs[2]/[3] scale on one GPU (1024 streams x 977 frames = 10^6 stereo frames),
    device-resident: (a) replicated streams give replicated bitstreams, (b) the 16 distinct streams
    match the reference, (c) the batch is reproducible after atde_reset()."""
    import torch
    tl.require_ref("test_full_size_properties (b)")
    S, F, C = 1024, 977, 2
    base = np.stack([tl.synth_rich(F, 1024, C, seed=9000 + s, kind=("mix", "tones", "steps")[s % 3]) if s % 4
                     else tl.synth_streams(1, F, 1024, C, seed=9000 + s)[0] for s in range(16)])
    d_base = torch.from_numpy(base).cuda()
    d_pcm = d_base.repeat(S // 16, 1, 1).contiguous()          # stream s == stream s % 16
    enc = ab.Encoder(ab.CODEC_ATRAC3, C, bitrate=kbit * 1024, lib=gpu_lib)
    fo, ub = enc.output_frames(F), enc.unit_bytes
    d_out = torch.empty((S, fo, ub), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    enc.encode_device(d_pcm.data_ptr(), S, F, d_out.data_ptr())
    enc.sync()
    out = d_out.cpu().numpy()
    assert np.array_equal(out[:16], out[16:32]) and np.array_equal(out[:16], out[-16:])
    assert (out.reshape(S // 16, 16, fo, ub) == out[:16]).all()
    for s in range(16):
        want = pc.oracle_at3(C, base[s].reshape(-1), kbit)
        bad = np.argwhere((out[s] != want).any(-1))
        assert bad.size == 0, (s, bad[:4, 0].tolist())
    enc.reset()
    d_out2 = torch.empty_like(d_out)
    enc.encode_device(d_pcm.data_ptr(), S, F, d_out2.data_ptr())
    enc.sync()
    assert torch.equal(d_out, d_out2)
    enc.close()

