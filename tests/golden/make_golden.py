#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference (/root/reference/src/duet) -- only runnable in the build container.

    python tests/golden/make_golden.py

The reference shells out to `samtools view` (sv_phasing_fn.py:25); a two-line PATH
shim that `cat`s the file stands in for it (the per-contig "BAMs" are SAM text).
Nothing here is imported by the test-suite; the tests read only the .json.gz files.

Outputs:
  kat_phase_info.json.gz   -- get_phase_info / predict_hp on hand-built and random calls
  e2e_<name>.json.gz       -- whole-stage cases: input files, joined reads, per-SV
                              feature trace, phased callset and phased_sv.vcf text
"""
import gzip
import json
import os
import stat
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from duet_b200 import synth  # noqa: E402


def install_shim():
    d = tempfile.mkdtemp(prefix="duet_shim_")
    p = os.path.join(d, "samtools")
    with open(p, "w") as f:
        f.write('#!/bin/bash\nexec cat "${@: -1}"\n')
    os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
    # `tabix --list-chroms <home>/snp_calling/pileup.vcf.gz` (read_file.py:15): the shim prints the
    # contig list stored next to the (non-existent) pileup file
    p = os.path.join(d, "tabix")
    with open(p, "w") as f:
        f.write('#!/bin/bash\nexec cat "$(dirname "${@: -1}")/contigs.txt"\n')
    os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
    os.environ["PATH"] = d + os.pathsep + os.environ["PATH"]


def py(o):
    if isinstance(o, dict):
        return {str(k): py(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [py(v) for v in o]
    if isinstance(o, np.integer):
        return int(o)
    if isinstance(o, np.floating):
        return float(o)
    return o


def dump(name, obj):
    path = os.path.join(HERE, name)
    raw = json.dumps(py(obj), sort_keys=True).encode()
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(raw)
    print(f"{name}: {len(raw)} B raw, {os.path.getsize(path)} B gz")


def run_case(ref, reads, pos, svread, refread, ps_num, oneps, note=""):
    call = {"svreadinfo": [list(r) for r in reads], "pos": pos, "svread": svread, "refread": refread}
    case = {"reads": reads, "pos": pos, "svread": svread, "refread": refread, "ps_num": ps_num,
            "oneps": sorted(oneps), "note": note}
    try:
        f = ref.get_phase_info(call, 2, ps_num, set(oneps))
        pred, ps = ref.predict_hp(call, ps_num, set(oneps))
        case["features"] = f
        case["pred"] = pred
        case["ps"] = ps
    except Exception as e:  # the exception type is part of the contract
        case["raises"] = type(e).__name__
    return case


def kat_cases(ref):
    R = lambda hap, ps, pc, nm="r": [nm, hap, ps, pc]
    cases = []
    add = lambda *a, **k: cases.append(run_case(ref, *a, **k))
    # SURVEY.md §4 table
    add([R(1, 200, 10), R(2, 100, 10), R(1, 100, 10), R(2, 200, 10)], 150, 4, 1, 2, {100, 200}, note="multi-PS tie")
    add([R(1, 300, 10), R(2, 100, 10), R(1, 300, 10)], 150, 3, 1, 2, {100}, note="PS not in oneps")
    add([R(1, 300, 10), R(2, 100, 10), R(1, 300, 10)], 1000, 3, 1, 2, {5000, 900, 1100}, note="none in oneps")
    for p in (1000, 10, 99999, 900, 901, 999, 1001, 1100, 1101):
        add([["x"]], p, 5, 0, 0, {900, 1100}, note="nearest")
    add([R(3, 100, 10)], 50, 3, 1, 2, {100}, note="bad HP class2")
    add([R(0, 100, 10)], 50, 3, 1, 2, {100}, note="HP 0 class2")
    add([R(3, 100, 10), R(2, 100, 10)], 50, 3, 1, 1, {100}, note="bad HP class1")
    add([R(1, 100, 8100), R(1, 100, 8101), R(2, 100, 0)], 50, 3, 3, 1, {100}, note="PC edge")
    add([R(1, 100, 9000), R(1, 100, 9000)], 50, 9, 1, 1, {777}, note="all PC>8100")
    add([R(1, 100, 50)] * 25, 50, 25, 0, 1, {100}, note="dead sv_num>=20")
    add([R(1, 100, 50)], 50, 0, 0, 1, {100}, note="zero division")
    add([["a"], ["b"]], 500, 4, 0, 0, {100, 900}, note="class0 emit")
    add([["a"], ["b"]], 500, 3, 0, 0, {100, 900}, note="class0 drop")
    # exact-threshold quotients
    for sv, rf in ((6, 19), (9, 1), (3, 1), (3, 7), (45, 55), (18, 7), (12, 38), (27, 3), (9, 11), (72, 28), (1, 0)):
        for reads in ([R(1, 100, 700), R(1, 100, 900), ["m"]],
                      [R(1, 100, 700), R(2, 100, 900), R(2, 100, 100)],
                      [R(1, 100, 8000), R(2, 100, 10), ["m"], ["n"]],
                      [R(2, 100, 700), ["m"], ["n"], ["o"], ["p"]]):
            add(reads, 120, sv, rf, 1, {100}, note="threshold grid c1")
        add([R(1, 100, 700), R(2, 200, 900), R(1, 200, 2500)], 120, sv, rf, 2, {100, 200}, note="threshold grid c2")
        add([R(1, 100, 10), R(2, 200, 3000), R(2, 200, 3000)] + [R(1, 300, 5)] * 6, 120, sv, rf, 2, {200}, note="hap0>=6")
        add([["m"]] * 3, 120, sv, rf, 0, {100, 200}, note="threshold grid c0")
    # seeded random calls around the thresholds
    rng = np.random.default_rng(7)
    for i in range(1500):
        ps_num = int(rng.integers(0, 3))
        n = int(rng.integers(1, 14))
        pss = [100, 200, 300, 4000][: int(rng.integers(1, 5))] if ps_num == 2 else [100]
        reads = []
        for j in range(n):
            if ps_num == 0 or rng.random() < 0.25:
                reads.append([f"m{j}"])
            else:
                pc = int(rng.choice([0, 5, 700, 2400, 2401, 4800, 8100, 8101, int(rng.integers(0, 12000))]))
                reads.append([f"r{j}", int(rng.integers(1, 3)), int(rng.choice(pss)), pc])
        oneps = set(int(x) for x in rng.choice([100, 200, 300, 4000, 77, 150], size=int(rng.integers(1, 5))))
        svread = int(rng.integers(1, 30))
        refread = int(rng.choice([0, 0, int(rng.integers(0, 41))]))
        cases.append(run_case(ref, reads, int(rng.integers(1, 5000)), svread, refread, ps_num, oneps, note="random"))
    return cases


def e2e_case(ref, sp, name, sample, dialect, svlen_thres=50, suppread_thres=2, mutate=None, all_ctgs=None):
    home = tempfile.mkdtemp(prefix="duet_golden_")
    synth.write_workdir(sample, home, dialect)
    if mutate:
        mutate(home)
    inc = all_ctgs is not None
    if inc:
        os.makedirs(home + "/snp_calling")
        with open(home + "/snp_calling/contigs.txt", "w") as f:
            f.write("".join(c + "\n" for c in all_ctgs))
    vcf = home + "/sv_calling/variants.vcf"
    sam_home = home + "/snp_phasing/"
    files = {}
    for sub in ("snp_phasing", "sv_calling") + (("snp_calling",) if inc else ()):
        for fn in sorted(os.listdir(os.path.join(home, sub))):
            with open(os.path.join(home, sub, fn)) as f:
                files[sub + "/" + fn] = f.read()
    # the join, straight from the reference
    read_hap = ref.read_hap_bam(sam_home, 1, inc)
    callinfo = ref.generate_callinfo(vcf, read_hap, inc)
    joined = [{"chrom": c["chrom"], "pos": c["pos"], "svlen": c["svlen"], "svtype": c["svtype"],
               "svread": c["svread"], "refread": c["refread"], "callgt": c["callgt"],
               "reads": [r[1:] for r in c["svreadinfo"]]} for c in callinfo]
    # per-SV feature trace: wrap (not modify) get_phase_info
    trace = []
    orig = ref.get_phase_info

    def spy(call, ps_sr=2, ps_num=1, oneps_set=""):
        f = orig(call, ps_sr, ps_num, oneps_set)
        trace.append({"chrom": call["chrom"], "pos": call["pos"], "ps_num": ps_num,
                      "f": {k: f[k] for k in ("hap1", "hap2", "hap0", "allhap", "nohap", "ps", "hap1_totsc",
                                              "hap2_totsc", "hapread_ratio", "sv_ratio", "hap1_avgsc",
                                              "hap2_avgsc", "totsc_ratio", "hap_avgsc_diff", "onehap_totsc")}})
        return f

    ref.get_phase_info = spy
    try:
        rows = ref.generate_phased_callset(vcf, sam_home, svlen_thres, suppread_thres, 1, inc)
    finally:
        ref.get_phase_info = orig
    sp.sv_phasing(home, svlen_thres, suppread_thres, 1, inc)
    with open(home + "/phased_sv.vcf") as f:
        out_text = f.read()
    dump(f"e2e_{name}.json.gz", {
        "name": name, "dialect": dialect, "svlen_thres": svlen_thres, "suppread_thres": suppread_thres,
        "include_all_ctgs": inc,
        "files": files, "joined": joined, "trace": trace, "rows": rows, "phased_sv_vcf": out_text})
    print(f"  {name}: {len(joined)} SVs, {len(rows)} phased rows")


def main():
    install_shim()
    import logging
    logging.disable(logging.CRITICAL)
    from duet import sv_phasing_fn as ref
    from duet import sv_phasing as sp

    dump("kat_phase_info.json.gz", kat_cases(ref))

    mk = synth.make_sample
    e2e_case(ref, sp, "cutesv_3ctg", mk(1, contigs=["1", "21", "X"], n_reads=3000, n_svs=260, bp_per_read=700, block_mean=1.5e5), "cutesv")
    e2e_case(ref, sp, "sniffles_chr", mk(2, contigs=["2", "22"], n_reads=2500, n_svs=200, chr_prefix=True,
                                        bp_per_read=700, block_mean=1.5e5), "sniffles")
    e2e_case(ref, sp, "svim_shuffled", mk(3, contigs=["5", "10", "Y"], n_reads=2500, n_svs=220, shuffle_vcf=True,
                                         bp_per_read=700, block_mean=1e5), "svim", svlen_thres=40, suppread_thres=3)
    e2e_case(ref, sp, "dense", mk(4, contigs=["20", "21"], n_reads=6000, n_svs=60, dense=True, bp_per_read=350, block_mean=3e5,
                                 empty_oneps_contig=None), "cutesv")

    # chrom ordering ('10' < 'chr1' as strings), duplicate QNAME last-wins, tie on (chrom,pos)
    def mixed_prefix(home):
        os.rename(home + "/snp_phasing/1.bam", home + "/snp_phasing/chr1.bam")
        p = home + "/sv_calling/variants.vcf"
        with open(p) as f:
            lines = f.readlines()
        out = []
        for ln in lines:
            if ln.startswith("1\t"):
                ln = "chr" + ln
            out.append(ln)
        # duplicate one record so two rows tie on (chrom, pos)
        body = [l for l in out if l.startswith("10\t")]
        out.append(body[0])
        with open(p, "w") as f:
            f.writelines(out)

    e2e_case(ref, sp, "mixed_prefix", mk(5, contigs=["1", "10"], n_reads=2000, n_svs=150, bp_per_read=700, block_mean=1e5,
                                        empty_oneps_contig=None), "cutesv", mutate=mixed_prefix)


    # -a / include_all_ctgs: the contig list comes from `tabix --list-chroms`; decoys, and a list naming
    # both '7' and 'chr7' (every 'chr7' record then belongs to two contigs and is phased twice)
    def rename_contigs(home):
        os.rename(home + "/snp_phasing/3.bam", home + "/snp_phasing/GL000192.1.bam")
        os.rename(home + "/snp_phasing/7.bam", home + "/snp_phasing/chr7.bam")
        p = home + "/sv_calling/variants.vcf"
        with open(p) as f:
            lines = f.readlines()
        out = []
        for ln in lines:
            ln = ln.replace("##contig=<ID=3,", "##contig=<ID=GL000192.1,").replace("##contig=<ID=7,", "##contig=<ID=chr7,")
            if ln.startswith("3\t"):
                ln = "GL000192.1" + ln[1:]
            elif ln.startswith("7\t"):
                ln = "chr" + ln
            out.append(ln)
        with open(p, "w") as f:
            f.writelines(out)

    e2e_case(ref, sp, "all_ctgs", mk(6, contigs=["2", "3", "7"], n_reads=2400, n_svs=180, bp_per_read=700, block_mean=1e5,
                                    empty_oneps_contig=None), "cutesv", mutate=rename_contigs,
             all_ctgs=["2", "GL000192.1", "7", "chr7", "MT"])


if __name__ == "__main__":
    main()
