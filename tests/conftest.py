import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "vq_golden.npz"))


@pytest.fixture(scope="session")
def patch_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "patch_golden.npz"))


@pytest.fixture(scope="session")
def patch_wide_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "patch_wide_golden.npz"))


@pytest.fixture(scope="session")
def bulk_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "bulk_golden.npz"))


@pytest.fixture(scope="session")
def manifest():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "manifest.json")) as f:
        return json.load(f)
