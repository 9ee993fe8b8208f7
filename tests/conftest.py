import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # On a box with a GPU the prebuilt oracle/_ref must have travelled with the snapshot: a parity test that cannot
    # reach the reference FAILS there instead of being skipped (atde_testlib.require_ref).  Opt out with
    # ATDE_REQUIRE_REF=0.
    import os
    if "ATDE_REQUIRE_REF" not in os.environ:
        try:
            import torch
            if torch.cuda.is_available():
                os.environ["ATDE_REQUIRE_REF"] = "1"
        except Exception:
            pass


@pytest.fixture(scope="session")
def ref():
    import atde_testlib as tl
    lib = tl.ref_lib()
    if lib is None:
        pytest.skip("oracle/_ref not built (reference sources absent on this box)")
    return lib


@pytest.fixture(scope="session")
def emu_lib():
    """The kernel sources compiled against the pthread CUDA shim (tests/cpuemu) — test tooling."""
    import atde_testlib as tl
    import atracdenc_b200 as ab
    tl.build_emu()
    return ab.load_library(tl.EMU_SO)


@pytest.fixture(scope="session")
def gpu_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import atracdenc_b200 as ab
    return ab.load_library()
