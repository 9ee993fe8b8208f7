"""CPU-only: the real kernel sources (atracdenc_b200/csrc/*.cu) compiled against the pthread CUDA
shim in tests/cpuemu, driven through the same C ABI, checked against the oracle at tiny sizes.
This is a test of the kernel LOGIC on a box without a GPU; the GPU suite repeats it on hardware."""
import parity_cases as pc


def test_golden_config1(emu_lib):
    pc.check_at1_golden_config1(emu_lib)


def test_golden_stereo_bursts(emu_lib):
    pc.check_at1_golden_stereo(emu_lib, max_frames=12)


def test_vs_oracle_default(emu_lib):
    pc.check_at1_vs_oracle(emu_lib, S=2, F=10, C=2)


def test_vs_oracle_forced_short_windows(emu_lib):
    pc.check_at1_vs_oracle(emu_lib, S=1, F=6, C=2, window_mask=7)
    pc.check_at1_vs_oracle(emu_lib, S=1, F=6, C=1, window_mask=2)


def test_vs_oracle_fixed_bfu(emu_lib):
    pc.check_at1_vs_oracle(emu_lib, S=1, F=6, C=2, bfu=3)


def test_stage_taps(emu_lib):
    pc.check_at1_stage_taps(emu_lib, S=1, F=9, C=2)


def test_batch_split_invariance(emu_lib):
    pc.check_at1_batch_split_invariance(emu_lib, S=2, F=9, C=2, cut=3)


def test_stream_independence(emu_lib):
    pc.check_at1_stream_independence(emu_lib, F=5)


def test_edge_inputs(emu_lib):
    pc.check_at1_edge_inputs(emu_lib)


def test_errors(emu_lib):
    pc.check_errors(emu_lib)
