"""TEST INFRASTRUCTURE: time the UNMODIFIED reference (oracle/_ref, installed by oracle/build_ref.py)
on a synthetic workload written out as the files it reads.  Used only by bench.py's CPU legs
(`--impl reference`, `cpu_baseline`) and by tests; never by the product path.

What is timed, all through the reference's own functions (sv_phasing_fn.py):
  stage_wall_s   generate_phased_callset(vcf, sam_home, 50, 2, thread, False): files in -> rows out  (:185-230)
  decode_s       read_hap_bam (:11-34, through the `samtools view` PATH shim) + parse_vcf (read_file.py:25-76)
  post-decode    generate_phased_callset with those two calls answered from memory (their own earlier return
                 values): the join (:46-48), classification, one-PS sets, predict_hp over every kept SV and
                 the sort -- the span the device path covers, and the figure `value` is quoted on
The reference's hot path is one Python thread (`thread` only reaches samtools): cores = 1.
"""
from __future__ import annotations

import logging
import os
import shutil
import sys
import tempfile
import time

from . import build_ref


def available() -> bool:
    return build_ref.available()


def _import_reference():
    if build_ref.SITE not in sys.path:
        sys.path.insert(0, build_ref.SITE)
    if build_ref.BIN not in os.environ.get("PATH", "").split(os.pathsep):
        os.environ["PATH"] = build_ref.BIN + os.pathsep + os.environ.get("PATH", "")
    from duet import read_file, sv_phasing_fn          # the unmodified modules
    assert os.path.realpath(sv_phasing_fn.__file__).startswith(os.path.realpath(build_ref.SITE)), sv_phasing_fn.__file__
    return read_file, sv_phasing_fn


class ReferenceRun:
    """One synthetic sample on disk + the reference's decoded objects, ready for timed post-decode steps."""

    def __init__(self, sample, *, dialect: str = "cutesv", tmp_root: str | None = None):
        from duet_b200 import synth
        self.read_file, self.fn = _import_reference()
        logging.getLogger().setLevel(logging.WARNING)       # the reference logs per contig
        self.home = tempfile.mkdtemp(prefix="duet_ref_run_", dir=tmp_root)
        t0 = time.perf_counter()
        synth.write_workdir(sample, self.home, dialect)
        self.write_s = time.perf_counter() - t0
        self.vcf = self.home + "/sv_calling/variants.vcf"
        self.sam_home = self.home + "/snp_phasing/"
        self.n_svs, self.n_joins, self.n_tagged = sample.n_svs, sample.n_joins, sample.n_tagged
        self.input_bytes = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(self.home) for f in fs)
        self.read_hap = None
        self.parsed = None

    def close(self):
        shutil.rmtree(self.home, ignore_errors=True)

    # -- whole stage, nothing patched ---------------------------------------------------------------
    def stage_wall(self) -> tuple[float, list]:
        t0 = time.perf_counter()
        rows = self.fn.generate_phased_callset(self.vcf, self.sam_home, 50, 2, 1, False)
        return time.perf_counter() - t0, rows

    # -- decode, by the reference's own functions -------------------------------------------------------
    def decode(self) -> dict:
        t0 = time.perf_counter()
        self.read_hap = self.fn.read_hap_bam(self.sam_home, 1, False)
        t1 = time.perf_counter()
        self.parsed = self.read_file.parse_vcf(self.vcf, False)
        t2 = time.perf_counter()
        return {"read_hap_bam_s": t1 - t0, "parse_vcf_s": t2 - t1, "decode_s": t2 - t0}

    # -- post-decode compute: the same function with its two decode calls answered from memory ---------------
    def post_decode_steps(self, n: int) -> tuple[list[float], list]:
        if self.read_hap is None:
            self.decode()
        # generate_callinfo overwrites column [13] of every record row (:46): each step gets fresh row lists
        fresh = [[[list(call) for call in ctg] for ctg in self.parsed] for _ in range(n)]
        orig_bam, orig_vcf = self.fn.read_hap_bam, self.fn.parse_vcf
        self.fn.read_hap_bam = lambda path, thread, inc: self.read_hap
        self.fn.parse_vcf = lambda path, inc: fresh.pop()
        times, rows = [], None
        try:
            for _ in range(n):
                t0 = time.perf_counter()
                rows = self.fn.generate_phased_callset(self.vcf, self.sam_home, 50, 2, 1, False)
                times.append(time.perf_counter() - t0)
        finally:
            self.fn.read_hap_bam, self.fn.parse_vcf = orig_bam, orig_vcf
        return times, rows
