"""CPU-only: the ATRAC3 kernel sources compiled against the pthread CUDA shim (tests/cpuemu), driven
through the C ABI, checked against the reference encoder (oracle/_ref) / golden fixtures at small
sizes.  The GPU suite repeats these on hardware at larger sizes."""
import parity_cases as pc


def test_golden_lp2(emu_lib):
    pc.check_at3_golden(emu_lib, "at3_lp2_stereo.npz", max_frames=10)


def test_golden_lp4(emu_lib):
    pc.check_at3_golden(emu_lib, "at3_lp4_js_stereo.npz", max_frames=10)


def test_vs_oracle_lp2(emu_lib):
    pc.check_at3_vs_oracle(emu_lib, S=3, F=9, C=2, kbit=0, seed=100)


def test_vs_oracle_lp4_joint_stereo(emu_lib):
    pc.check_at3_vs_oracle(emu_lib, S=3, F=9, C=2, kbit=64, seed=200)


def test_vs_oracle_mono_lp2(emu_lib):
    pc.check_at3_vs_oracle(emu_lib, S=2, F=7, C=1, kbit=0, seed=300)


def test_vs_oracle_mono_in_joint_stereo_container(emu_lib):
    """Mono input at LP4 / 94 kbit: the reference writes an empty second element and gives the first one all the
    bytes the second cannot use (atrac3denc.cpp:843-849, atrac3_bitstream.cpp:750-752)."""
    pc.check_at3_vs_oracle(emu_lib, S=3, F=8, C=1, kbit=64, seed=320)
    pc.check_at3_vs_oracle(emu_lib, S=1, F=6, C=1, kbit=90, seed=330, kinds=("steps",))


def test_vs_oracle_flags(emu_lib):
    pc.check_at3_vs_oracle(emu_lib, S=1, F=7, C=2, kbit=0, seed=400, kinds=("mix",), no_gain=1)
    pc.check_at3_vs_oracle(emu_lib, S=1, F=7, C=2, kbit=0, seed=401, kinds=("tones",), no_tonal=1)


def test_stage_taps(emu_lib):
    pc.check_at3_stage_taps(emu_lib, C=2, F=9, kbit=0)
    pc.check_at3_stage_taps(emu_lib, C=2, F=9, kbit=64, kind="steps")


def test_main_loop_view(emu_lib):
    pc.check_at3_main_loop(emu_lib, C=2, kbit=0, seconds=0.3)


def test_batch_split_invariance(emu_lib):
    pc.check_at3_batch_split_invariance(emu_lib, S=2, F=9, C=2, kbit=0, cuts=(3, 2))
    pc.check_at3_batch_split_invariance(emu_lib, S=1, F=8, C=2, kbit=64, cuts=(1, 3))


def test_stream_independence(emu_lib):
    pc.check_at3_stream_independence(emu_lib, F=5)


def test_edge_inputs(emu_lib):
    pc.check_at3_edge_inputs(emu_lib)


def test_errors(emu_lib):
    pc.check_at3_errors(emu_lib)


def test_gain_trace_taps(emu_lib):
    pc.check_at3_gain_trace_taps(emu_lib, S=2, F=6, C=2)
    pc.check_at3_gain_trace_taps(emu_lib, S=1, F=5, C=2, kbit=64, seed=2200)      # joint stereo: the trace is over M/S
