"""CPU-only: the ATRAC3plus stage kernels compiled against the pthread CUDA shim (tests/cpuemu),
checked against the reference's taps (oracle/_ref).  The GPU suite repeats these on hardware."""
import parity_cases as pc


def test_pqf(emu_lib):
    pc.check_at3p_pqf(emu_lib, S=2, F=3, C=2)
    pc.check_at3p_pqf(emu_lib, S=1, F=2, C=1, seed=905)


def test_mdct(emu_lib):
    pc.check_at3p_mdct(emu_lib, S=2, F=4, C=2)
    pc.check_at3p_mdct(emu_lib, S=1, F=3, C=1, seed=915)


def test_pack(emu_lib):
    pc.check_at3p_pack(emu_lib, S=3, F=5, C=2)
    pc.check_at3p_pack(emu_lib, S=1, F=4, C=1, seed=925)
    pc.check_at3p_pack(emu_lib, S=1, F=4, C=2, seed=930, loud=True)


def test_pack_random_spectra(emu_lib):
    nq = pc.check_at3p_pack_random(emu_lib, U=10, C=2)
    assert nq == 0 or min(nq) < 28
    pc.check_at3p_pack_random(emu_lib, U=6, C=1, seed=950)


def test_tone_filter(emu_lib):
    pc.check_at3p_tone_filter(emu_lib, S=3, F=6, C=2)
    pc.check_at3p_tone_filter(emu_lib, S=1, F=5, C=1, seed=965)


def test_chain_after_gha(emu_lib):
    pc.check_at3p_chain_after_gha(emu_lib, S=2, F=5, C=2)


def test_trig_replicas_match_libm(emu_lib):
    """glibc_trig.cuh (compiled for the host by the emulator build) against the live libm."""
    pc.check_trig_replicas(emu_lib, n=200000)


def test_gha_search(emu_lib):
    pc.check_at3p_gha(emu_lib, S=2, F=3, C=2)


def test_full_chain(emu_lib):
    pc.check_at3p_full_chain(emu_lib, S=2, F=5, C=2)
    pc.check_at3p_full_chain(emu_lib, S=1, F=4, C=1, seed=995)


def test_encoder_vs_oracle(emu_lib):
    pc.check_at3p_vs_oracle(emu_lib, S=3, F=6, C=2)
    pc.check_at3p_vs_oracle(emu_lib, S=1, F=5, C=1, seed=1210)


def test_encoder_batch_split_invariance(emu_lib):
    pc.check_at3p_batch_split_invariance(emu_lib, S=2, F=9, C=2)
    pc.check_at3p_batch_split_invariance(emu_lib, S=1, F=7, C=1, cuts=(2, 1, 1), seed=1310)


def test_host_chunking(emu_lib):
    import atracdenc_b200 as ab
    pc.check_host_chunking(emu_lib, ab.CODEC_ATRAC1, S=7, F=3)
    pc.check_host_chunking(emu_lib, ab.CODEC_ATRAC1, S=27, F=2, variants=(None, "8", "8/i16"))
    pc.check_host_chunking(emu_lib, ab.CODEC_ATRAC3, S=7, F=2)
    pc.check_host_chunking(emu_lib, ab.CODEC_ATRAC3PLUS, S=6, F=2)


def test_i16_ingest(emu_lib):
    import atracdenc_b200 as ab
    pc.check_i16_ingest(emu_lib, ab.CODEC_ATRAC1, S=3, F=5)
    pc.check_i16_ingest(emu_lib, ab.CODEC_ATRAC1, C=1, S=2, F=3, seed=1810)
    pc.check_i16_ingest(emu_lib, ab.CODEC_ATRAC3, S=3, F=4)
    pc.check_i16_ingest(emu_lib, ab.CODEC_ATRAC3PLUS, S=2, F=3)


def test_gha_debug_masks(emu_lib):
    pc.check_at3p_gha_masks(emu_lib, masks=(0, 3, 5, 6), S=2, F=5)


def test_edge_inputs(emu_lib):
    pc.check_at3p_edge_inputs(emu_lib, C=2, F=4)
    pc.check_at3p_edge_inputs(emu_lib, C=1, F=4)


def test_mono_and_stream_independence(emu_lib):
    pc.check_at3p_vs_oracle(emu_lib, S=2, F=4, C=1, seed=2200)
    pc.check_at3p_stream_independence(emu_lib, F=4)


def test_golden(emu_lib):
    pc.check_at3p_golden(emu_lib, "at3p_stereo.npz", max_frames=5)
    pc.check_at3p_golden(emu_lib, "at3p_mono.npz", max_frames=5)
