#!/usr/bin/env python
"""Wall time of the whole drop-in stage -- duet_b200.sv_phasing.sv_phasing(home, ...) -- on the C2
workload written out as the files the reference reads (per-contig SAM text + the cuteSV VCF), next to
the oracle port of the reference run on the same files with the same one Python thread.
    python tools/stage_wall_time.py [c2|c1] [threads]
Prints one JSON line.  Needs a GPU (the product stage has no CPU fallback)."""
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import make_sample  # noqa: E402
from duet_b200 import sv_phasing, sv_phasing_fn, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
threads = int(sys.argv[2]) if len(sys.argv) > 2 else min(8, os.cpu_count() or 1)
sample = make_sample(wl, 0)
home = tempfile.mkdtemp(prefix="duet_stage_")
try:
    t0 = time.perf_counter()
    synth.write_workdir(sample, home)
    write_s = time.perf_counter() - t0
    sam_mb = sum(os.path.getsize(os.path.join(home, "snp_phasing", f)) for f in os.listdir(os.path.join(home, "snp_phasing"))) / 1e6
    vcf_mb = os.path.getsize(os.path.join(home, "sv_calling", "variants.vcf")) / 1e6
    sv_phasing_fn.get_engine()                                    # context creation is not part of the stage
    runs = []
    for _ in range(3):
        t0 = time.perf_counter()
        sv_phasing.sv_phasing(home, 50, 2, threads, False)
        runs.append((time.perf_counter() - t0, dict(sv_phasing_fn.last_timings)))
    best, tm = min(runs, key=lambda r: r[0])
    out = open(os.path.join(home, "phased_sv.vcf")).read()
    line = {"workload": wl, "threads": threads, "sam_text_mb": round(sam_mb, 1), "vcf_mb": round(vcf_mb, 1),
            "stage_wall_s": round(best, 3), "host_decode_s": round(tm["host_decode_s"], 3),
            "device_call_s": round(tm["device_call_s"], 4), "rows_s": round(tm["rows_s"], 3),
            "rows_written": out.count("\n") - sum(1 for l in out.splitlines() if l.startswith("#")),
            "workdir_write_s": round(write_s, 1)}
    if "--port" in sys.argv:                                      # the oracle port of the reference on the same files
        from oracle import ref_port
        shim = tempfile.mkdtemp(prefix="duet_shim_")
        with open(os.path.join(shim, "samtools"), "w") as f:
            f.write('#!/bin/bash\nexec cat "${@: -1}"\n')
        os.chmod(os.path.join(shim, "samtools"), 0o755)
        os.environ["PATH"] = shim + os.pathsep + os.environ["PATH"]
        t0 = time.perf_counter()
        rows = ref_port.generate_phased_callset(home + "/sv_calling/variants.vcf", home + "/snp_phasing/", 50, 2, 1, False)
        line["port_wall_s"] = round(time.perf_counter() - t0, 2)
        line["port_rows"] = len(rows)
    print(json.dumps(line))
finally:
    shutil.rmtree(home, ignore_errors=True)
