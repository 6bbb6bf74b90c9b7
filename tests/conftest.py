import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Product library and oracle are built before any test runs."""
    from nxsearch_b200 import _build
    import _oracle

    _build.build()
    _oracle.build_port()
    _oracle.build_ref()


@pytest.fixture(scope="session")
def c1_corpus():
    """BASELINE config 1: 10k docs, 50k-term Zipf vocabulary."""
    from nxsearch_b200 import tools

    return tools.Corpus.generate(10_000, 50_000)


@pytest.fixture(scope="session")
def c1_oracle(c1_corpus):
    import _oracle

    return _oracle.OracleIndex(c1_corpus)
