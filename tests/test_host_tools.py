"""CPU-only checks of two pieces of kernel logic that parity depends on, compiled for the host from
the kernel sources (tests/cpuemu shim):
  * stdsort_dev.cuh  vs the real libstdc++ std::sort / std::partial_sort (tie order included)
  * at3_pack.cu's warp-cooperative quantiser (compute_units: fast walk with early termination and the
    exact fallback, CLC/VLC costs) vs the reference's own QuantMantisas linked from oracle/_ref."""
import subprocess
from pathlib import Path

import pytest

import atde_testlib as tl

ROOT = Path(__file__).resolve().parent.parent
FLAGS = ["-O1", "-std=c++17", "-DATDE_CPU_EMU", f"-I{ROOT}/tests/cpuemu", f"-I{ROOT}/atracdenc_b200/csrc",
         "-pthread", "-Wno-attributes", "-ffp-contract=off"]


def _build(tmp_path, src, extra=()):
    exe = tmp_path / Path(src).stem
    subprocess.check_call(["g++", *FLAGS, str(ROOT / "tests" / "tools" / src), str(ROOT / "tests" / "cpuemu" / "cuda_emu.cpp"),
                           "-o", str(exe), *extra])
    return exe


def test_stdsort_replica_matches_libstdcxx(tmp_path):
    exe = _build(tmp_path, "stdsort_check.cpp")
    out = subprocess.check_output([str(exe), "60000"], text=True)
    assert "mismatches=0" in out, out


def test_quantiser_matches_reference(tmp_path):
    if tl.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    ref_dir = ROOT / "oracle" / "_ref"
    exe = _build(tmp_path, "quant_check.cpp", [f"-L{ref_dir}", "-latde_ref", f"-Wl,-rpath,{ref_dir}"])
    out = subprocess.check_output([str(exe), "500"], text=True)
    assert "mismatches=0" in out, out
