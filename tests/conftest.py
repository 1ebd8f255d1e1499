import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class G:
        def __getattr__(self, name):
            return np.load(os.path.join(GOLDEN, name + ".npz"))
    return G()


class _ParityLog(dict):
    """Deviations the GPU parity tests measure (not only assert): dumped as JSON at session end when SAPCU_PARITY_JSON
    names a file -- the archived copy lives under profiles/ (r02_parity.json)."""

    def record(self, test, **values):
        self.setdefault(test, {}).update(values)


_PARITY = _ParityLog()


@pytest.fixture(scope="session")
def parity_log():
    return _PARITY


def pytest_sessionfinish(session, exitstatus):
    path = os.environ.get("SAPCU_PARITY_JSON")
    if path and _PARITY:
        import json
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "w") as f:
            json.dump(_PARITY, f, indent=1, sort_keys=True)
