import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than ~10 s on CPU")


@pytest.fixture(scope="session")
def orc():
    from oracle_lib import Oracle, build_oracle

    build_oracle()
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    import json

    return json.loads((Path(__file__).resolve().parent / "golden" / "golden.json").read_text())


@pytest.fixture(scope="session")
def ref_fast():
    from oracle_lib import RefShim

    if not RefShim.available("fast"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return RefShim("fast")


@pytest.fixture(scope="session")
def ref_strict():
    from oracle_lib import RefShim

    if not RefShim.available("strict"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return RefShim("strict")
