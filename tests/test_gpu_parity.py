"""GPU parity suite: libatde_b200.so through the C ABI on a real B200 vs the oracle."""
import numpy as np
import pytest

import atde_testlib as tl
import atracdenc_b200 as ab
import parity_cases as pc

pytestmark = pytest.mark.gpu


def test_golden_config1(gpu_lib):
    pc.check_at1_golden_config1(gpu_lib)


def test_golden_stereo_bursts(gpu_lib):
    pc.check_at1_golden_stereo(gpu_lib)


@pytest.mark.parametrize("S,F,C", [(4, 33, 2), (3, 64, 1), (16, 16, 2)])
def test_vs_oracle_default(gpu_lib, S, F, C):
    pc.check_at1_vs_oracle(gpu_lib, S=S, F=F, C=C, seed=S * 100 + F)


@pytest.mark.parametrize("mask", [1, 2, 4, 7, 0])
def test_vs_oracle_forced_windows(gpu_lib, mask):
    pc.check_at1_vs_oracle(gpu_lib, S=2, F=17, C=2, window_mask=mask)


@pytest.mark.parametrize("bfu", [1, 3, 8])
def test_vs_oracle_fixed_bfu(gpu_lib, bfu):
    pc.check_at1_vs_oracle(gpu_lib, S=2, F=17, C=2, bfu=bfu)


def test_stage_taps(gpu_lib):
    pc.check_at1_stage_taps(gpu_lib, S=3, F=40, C=2)


def test_batch_split_invariance(gpu_lib):
    pc.check_at1_batch_split_invariance(gpu_lib, S=5, F=41, C=2, cut=13)


def test_stream_independence(gpu_lib):
    pc.check_at1_stream_independence(gpu_lib, F=19)


def test_edge_inputs(gpu_lib):
    pc.check_at1_edge_inputs(gpu_lib)


def test_errors(gpu_lib):
    pc.check_errors(gpu_lib)


def test_larger_batch_sampled_against_oracle(gpu_lib):
    """256 streams x 120 frames stereo (61 k channel-frames): every stream checked against the oracle."""
    S, F, C = 256, 120, 2
    pcm = tl.synth_streams(S, F, 512, C, seed=77)
    enc = ab.Encoder(ab.CODEC_ATRAC1, C, lib=gpu_lib)
    out = enc.encode(pcm, S)
    enc.close()
    for s in range(S):
        want, _ = pc.oracle_at1(C, pcm[s].reshape(-1))
        assert np.array_equal(out[s], want), s


def test_full_size_properties(gpu_lib):
    """BASELINE.json configs[1] scale (1024 streams x 977 frames stereo = 10^6 frames) without an
    oracle pass: (a) the device-resident and host entry points agree byte for byte, (b) the batch
    is reproducible, (c) replicated streams give replicated bitstreams, (d) a sample of streams
    matches the oracle."""
    import torch
    S, F, C = 1024, 977, 2
    base = tl.synth_streams(16, F, 512, C, seed=123)
    pcm = np.tile(base, (S // 16, 1, 1))                    # stream s == stream s % 16
    enc = ab.Encoder(ab.CODEC_ATRAC1, C, lib=gpu_lib)
    out = enc.encode(pcm, S)
    # (c) replicas
    assert np.array_equal(out[:16], out[16:32]) and np.array_equal(out[:16], out[-16:])
    # (d) oracle on the 16 distinct streams
    for s in range(16):
        want, _ = pc.oracle_at1(C, base[s].reshape(-1))
        assert np.array_equal(out[s], want), s
    # (a)+(b) device path
    enc.reset()
    d_pcm = torch.from_numpy(pcm).cuda()
    d_out = torch.empty((S, F, C, 212), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    enc.encode_device(d_pcm.data_ptr(), S, F, d_out.data_ptr())
    enc.sync()
    assert np.array_equal(d_out.cpu().numpy(), out)
    enc.close()
