import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_transforms():
    import numpy as np
    z = np.load(os.path.join(GOLDEN, "transforms.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split("/", 1)
        cases.setdefault(name, {})[field] = z[key]
    for c in cases.values():
        c["kind"] = str(c["kind"])
        c["kw"] = eval(str(c["kw"]), {"__builtins__": {}}, {"True": True, "False": False})
    return cases
