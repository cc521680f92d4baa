import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import __graft_entry__ as g

        g.build()
        return g.load_package()._lib.load().p2p_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def pkg():
    """The product package (requires the built library; no GPU needed to import)."""
    import __graft_entry__ as g

    g.build()
    return g.load_package()


@pytest.fixture(scope="session")
def proj(pkg):
    """A device context on cuda:0.  GPU tests fail loudly (not skip) when no device exists."""
    p = pkg.Projector(0, n_slots=4)
    yield p
    p.close()


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
