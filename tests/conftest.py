import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for _p in (str(ROOT), str(Path(__file__).resolve().parent)):      # repo root + tests/ (helper modules)
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    from ball_action_spotting_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def oracle_sd():
    from oracle import mds_oracle as O
    return O.make_state_dict(O.ModelConfig(), seed=1234)
