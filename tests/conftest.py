import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference tree under /root/reference (build container only)")


def golden_names():
    """Episode fixtures recorded from the reference simulator (t*.npz, simple_rail_*.npz); other fixtures in the
    directory (MotionCheck cases, policy outputs, saved levels) have their own tests."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and (f.startswith("t") or f.startswith("simple_rail")))


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    cache = {}

    def load(name):
        if name not in cache:
            with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
                cache[name] = {k: z[k] for k in z.files}
        return cache[name]
    return load
