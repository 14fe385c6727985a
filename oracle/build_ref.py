#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- recipe that installs the UNMODIFIED reference into oracle/_ref/.

    python oracle/build_ref.py            # needs /root/reference (the build container)

The reference's hot path is pure Python (numpy only), so "building" it is a pip install of its own
setup.py into a private target directory:

    oracle/_ref/site/duet/*.py     the package, byte-identical to /root/reference/src/duet/*.py
    oracle/_ref/bin/samtools       PATH shim for `samtools view -@N <file>` (sv_phasing_fn.py:25): cats the
                                   file -- the per-contig "BAMs" of the synthetic workloads are SAM text
    oracle/_ref/bin/tabix          PATH shim for `tabix --list-chroms` (read_file.py:15)
    oracle/_ref/MANIFEST.json      sha256 of every installed module next to the sha256 of its source

oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so it
travels to the GPU box, where /root/reference does not exist.  Only tests/, bench.py's CPU legs and
__graft_entry__.build() touch it; nothing under duet_b200/ does.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import stat
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
DEST = os.path.join(HERE, "_ref")
SITE = os.path.join(DEST, "site")
BIN = os.path.join(DEST, "bin")

SHIMS = {
    "samtools": '#!/bin/bash\n# stands in for `samtools view -@N <file>`: the synthetic per-contig BAMs are SAM text\nexec cat "${@: -1}"\n',
    "tabix": '#!/bin/bash\n# stands in for `tabix --list-chroms <home>/snp_calling/pileup.vcf.gz`\nexec cat "$(dirname "${@: -1}")/contigs.txt"\n',
}


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def available() -> bool:
    return os.path.exists(os.path.join(SITE, "duet", "sv_phasing_fn.py")) and os.path.exists(os.path.join(BIN, "samtools"))


def build_ref(force: bool = False) -> bool:
    """Returns True when oracle/_ref holds the reference afterwards."""
    if available() and not force:
        return True
    if not os.path.isdir(os.path.join(REF_SRC, "src", "duet")):
        return available()
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(SITE)
    os.makedirs(BIN)
    with tempfile.TemporaryDirectory(prefix="duet_ref_src_") as tmp:
        src = os.path.join(tmp, "reference")                  # setup.py writes build/ and egg-info: use a copy
        shutil.copytree(REF_SRC, src)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--no-compile", "--target", SITE, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or not os.path.exists(os.path.join(SITE, "duet", "sv_phasing_fn.py")):
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("pip install of the reference into oracle/_ref/site failed")
    manifest = {}
    for fn in sorted(os.listdir(os.path.join(SITE, "duet"))):
        if fn.endswith(".py"):
            got, want = _sha(os.path.join(SITE, "duet", fn)), _sha(os.path.join(REF_SRC, "src", "duet", fn))
            if got != want:
                raise RuntimeError(f"installed {fn} differs from the reference source")
            manifest[fn] = got
    for name, text in SHIMS.items():
        p = os.path.join(BIN, name)
        with open(p, "w") as f:
            f.write(text)
        os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC | stat.S_IXGRP | stat.S_IXOTH)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC + "/src/duet", "installed_by": "pip install --target (setup.py of the reference)",
                   "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = build_ref(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference here)")
