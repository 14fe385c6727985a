import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name), "rb") as f:
        return json.loads(f.read().decode())


def golden_e2e_names():
    return sorted(fn[4:-8] for fn in os.listdir(GOLDEN) if fn.startswith("e2e_") and fn.endswith(".json.gz"))


def materialise(case, home):
    """Write a golden case's input files under `home`."""
    for rel, text in case["files"].items():
        p = os.path.join(home, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(text)
    return home


@pytest.fixture
def golden_workdir(tmp_path):
    def make(name):
        case = load_golden(f"e2e_{name}.json.gz")
        materialise(case, str(tmp_path))
        return case, str(tmp_path)
    return make
