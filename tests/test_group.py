"""atde_create_group (include/atde_b200.h): several encoders in one process, a batch sharded by stream over them.
The result must equal one encoder over the whole batch (streams are independent), across calls too (carried state)."""
import numpy as np
import pytest

import atde_testlib as tl
import atracdenc_b200 as ab


def check_group(lib, devices, codec, step, S=5, F=6, kbit=0):
    pcm = tl.synth_streams(S, F, step, 2, seed=77)
    one = ab.Encoder(codec, 2, bitrate=kbit * 1024, lib=lib, device=devices[0])
    want_a = one.encode(pcm[:, :2 * step], S)
    want_b = one.encode(pcm[:, 2 * step:], S)
    one.close()
    grp = ab.EncoderGroup(codec, 2, devices, bitrate=kbit * 1024, lib=lib)
    assert grp.size() == len(devices)
    got_a, sizes_a = grp.encode(pcm[:, :2 * step], S, want_sizes=True)
    got_b = grp.encode(pcm[:, 2 * step:], S)                    # streams continue on their member
    assert np.array_equal(got_a, want_a) and np.array_equal(got_b, want_b)
    assert (sizes_a > 0).all()
    grp.reset()
    i16 = np.rint(pcm * 32768).astype(np.int16)
    assert np.array_equal(grp.encode(i16[:, :2 * step], S), want_a)
    with pytest.raises(ab.AtdeError):
        grp.encode(pcm[:1], 1)                                  # fewer streams than members
    grp.close()


def test_group_emulated(emu_lib):
    check_group(emu_lib, [0, 0, 0], ab.CODEC_ATRAC1, 512, S=5, F=5)
    check_group(emu_lib, [0, 0], ab.CODEC_ATRAC3, 1024, S=3, F=5, kbit=64)


@pytest.mark.gpu
def test_group_gpu(gpu_lib):
    import torch
    n = torch.cuda.device_count()
    devices = list(range(n)) if n > 1 else [0, 0, 0]
    check_group(gpu_lib, devices, ab.CODEC_ATRAC1, 512, S=37, F=40)
    check_group(gpu_lib, devices, ab.CODEC_ATRAC3, 1024, S=23, F=30)
    check_group(gpu_lib, devices[:2], ab.CODEC_ATRAC3PLUS, 2048, S=5, F=8)
