import gzip
import json
import os
import sys

import stat
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def tabix_shim():
    """`tabix --list-chroms <home>/snp_calling/pileup.vcf.gz` (reference read_file.py:15) is only run with
    include_all_ctgs=True; the golden case for it stores the contig list next to the (absent) pileup file
    and this shim prints it -- the same shim tests/golden/make_golden.py gave the reference."""
    d = tempfile.mkdtemp(prefix="duet_shim_")
    p = os.path.join(d, "tabix")
    with open(p, "w") as f:
        f.write('#!/bin/bash\nexec cat "$(dirname "${@: -1}")/contigs.txt"\n')
    os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
    old = os.environ["PATH"]
    os.environ["PATH"] = d + os.pathsep + old
    yield
    os.environ["PATH"] = old


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
