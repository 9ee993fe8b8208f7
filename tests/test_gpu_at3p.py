"""GPU parity suite, ATRAC3plus stage kernels (PQF analysis, MDCT-256 x16) against the
reference's taps (oracle/_ref)."""
import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu


def test_pqf(gpu_lib):
    assert pc.check_at3p_pqf(gpu_lib, S=8, F=24, C=2) in (0, 8)
    pc.check_at3p_pqf(gpu_lib, S=3, F=10, C=1, seed=905)


def test_mdct(gpu_lib):
    pc.check_at3p_mdct(gpu_lib, S=6, F=20, C=2)
    pc.check_at3p_mdct(gpu_lib, S=2, F=9, C=1, seed=915)


def test_pack(gpu_lib):
    pc.check_at3p_pack(gpu_lib, S=6, F=20, C=2)
    pc.check_at3p_pack(gpu_lib, S=2, F=9, C=1, seed=925)
    pc.check_at3p_pack(gpu_lib, S=2, F=9, C=2, seed=930, loud=True)


def test_pack_random_spectra(gpu_lib):
    nq = pc.check_at3p_pack_random(gpu_lib, U=40, C=2)
    assert nq == 0 or min(nq) < 28
    pc.check_at3p_pack_random(gpu_lib, U=20, C=1, seed=950)


def test_tone_filter(gpu_lib):
    pc.check_at3p_tone_filter(gpu_lib, S=8, F=24, C=2)
    pc.check_at3p_tone_filter(gpu_lib, S=3, F=12, C=1, seed=965)


def test_chain_after_gha(gpu_lib):
    pc.check_at3p_chain_after_gha(gpu_lib, S=6, F=20, C=2)
    pc.check_at3p_chain_after_gha(gpu_lib, S=2, F=10, C=1, seed=975)


def test_gha_search(gpu_lib):
    pc.check_at3p_gha(gpu_lib, S=8, F=16, C=2)
    pc.check_at3p_gha(gpu_lib, S=3, F=10, C=1, seed=985)


def test_full_chain(gpu_lib):
    pc.check_at3p_full_chain(gpu_lib, S=6, F=16, C=2)
    pc.check_at3p_full_chain(gpu_lib, S=2, F=8, C=1, seed=995)


def test_encoder_vs_oracle(gpu_lib):
    assert pc.check_at3p_vs_oracle(gpu_lib, S=8, F=24, C=2) in (0, 8)
    pc.check_at3p_vs_oracle(gpu_lib, S=3, F=12, C=1, seed=1210)


def test_encoder_batch_split_invariance(gpu_lib):
    pc.check_at3p_batch_split_invariance(gpu_lib, S=4, F=20, C=2, cuts=(1, 7, 4))
    pc.check_at3p_batch_split_invariance(gpu_lib, S=2, F=9, C=1, cuts=(2, 1, 3), seed=1310)


def test_host_chunking(gpu_lib):
    import atracdenc_b200 as ab
    pc.check_host_chunking(gpu_lib, ab.CODEC_ATRAC1, S=11, F=4)
    pc.check_host_chunking(gpu_lib, ab.CODEC_ATRAC3, S=11, F=3)
    pc.check_host_chunking(gpu_lib, ab.CODEC_ATRAC3PLUS, S=10, F=2)


def test_i16_ingest(gpu_lib):
    import atracdenc_b200 as ab
    pc.check_i16_ingest(gpu_lib, ab.CODEC_ATRAC1, S=3, F=5)
    pc.check_i16_ingest(gpu_lib, ab.CODEC_ATRAC1, C=1, S=2, F=3, seed=1810)
    pc.check_i16_ingest(gpu_lib, ab.CODEC_ATRAC3, S=3, F=4)
    pc.check_i16_ingest(gpu_lib, ab.CODEC_ATRAC3PLUS, S=2, F=3)


def test_gha_debug_masks(gpu_lib):
    pc.check_at3p_gha_masks(gpu_lib, S=3, F=9)
