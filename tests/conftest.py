import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
GOLD = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def ptb():
    import ptb200
    return ptb200


@pytest.fixture(scope="session")
def oracle():
    """ctypes view of oracle/_build/libpt_oracle.so (the CPU restatement = the checker)."""
    import _oracle
    return _oracle.load()


@pytest.fixture(scope="session")
def duck(ptb):
    return ptb.load_scene_file(GOLD / "cornell_duck.ptscene.gz")


@pytest.fixture(scope="session")
def box(ptb):
    return ptb.load_scene_file(GOLD / "cornell_box.ptscene.gz")


@pytest.fixture(scope="session")
def core_lib(ptb):
    """libptcore.so, built in-tree if absent (nvcc cross-compiles without a GPU)."""
    if not ptb.LIB_PATH.exists():
        sys.path.insert(0, str(ROOT))
        import __graft_entry__
        __graft_entry__.build()
    return ptb.load_library()


@pytest.fixture()
def tracer(ptb, core_lib):
    t = ptb.PathTracer(0)
    yield t
    t.close()
