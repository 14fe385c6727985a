"""CPU oracle for the Duet sv_phasing hot path -- TEST INFRASTRUCTURE ONLY.

A plain-Python restatement of the reference algorithm, written from the
reference's behaviour (not its text) so it can travel to the GPU box where
/root/reference does not exist.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the
product (duet_b200/) never does.

Parity pinning: the reference has no golden vectors of its own,
so this port is pinned by running the UNMODIFIED reference here
(tests/golden/make_golden.py imports /root/reference/src and records its
outputs) -- the committed fixtures under tests/golden/ are the reference's own
outputs and tests/test_oracle_golden.py checks this port against all of them.

Reference citations are to /root/reference/src/duet/.
"""
from __future__ import annotations

import os
import shlex
import shutil
import subprocess
from bisect import bisect_left

PC_MAX = 8100  # sv_phasing_fn.py:76,88,201

GT_TEXT = {1: "1|0", 2: "0|1", 3: "1|1"}  # sv_phasing_fn.py:217-222


# ---------------------------------------------------------------------------
# contig list (read_file.py:6-16)
# ---------------------------------------------------------------------------

def contig_names(include_all_ctgs: bool, home: str) -> list[str]:
    if include_all_ctgs:
        vcf = home + "/snp_calling/pileup.vcf.gz"
        txt = subprocess.check_output(shlex.split("tabix --list-chroms " + vcf)).decode("ascii")
        return txt.split("\n")[:-1]
    return [str(n) for n in range(1, 23)] + ["X", "Y"]


# ---------------------------------------------------------------------------
# haplotag table (sv_phasing_fn.py:11-34)
# ---------------------------------------------------------------------------

def _sam_text(path: str, thread: int) -> str:
    """`samtools view` text of one per-contig BAM (sv_phasing_fn.py:25).  When the
    file is not gzip/BGZF it already is SAM text and is read directly (the GPU box
    has no samtools; test inputs are SAM text)."""
    with open(path, "rb") as fh:
        magic = fh.read(2)
    if magic == b"\x1f\x8b" or shutil.which("samtools"):
        return subprocess.check_output(
            shlex.split("samtools view -@" + str(thread) + " " + path)).decode("ascii")
    with open(path, "rb") as fh:
        return fh.read().decode("ascii")


def haplotag_tables(sam_home: str, thread: int, include_all_ctgs: bool) -> list[dict]:
    """One dict per contig: QNAME -> (hap, ps, pc).  A row is kept when its
    second-to-last field contains 'PC:i:'; HP/PC/PS are taken POSITIONALLY from the
    last three fields (:28-29); later rows overwrite earlier ones (:29); text after
    the final newline is dropped (:25 `[:-1]`); a missing BAM leaves an empty dict
    (:19-24)."""
    names = contig_names(include_all_ctgs, sam_home[:len(sam_home) - 13])
    tables = []
    for ctg in names:
        table = {}
        tables.append(table)
        path = None
        for cand in (sam_home + "chr" + ctg + ".bam", sam_home + ctg + ".bam"):
            if os.path.exists(cand):
                path = cand
                break
        if path is None:
            continue
        for line in _sam_text(path, thread).split("\n")[:-1]:
            f = line.split()
            if "PC:i:" in f[-2]:
                table[f[0]] = (int(f[-3][5:]), int(f[-1][5:]), int(f[-2][5:]))
    return tables


# ---------------------------------------------------------------------------
# SV records (read_file.py:18-76)
# ---------------------------------------------------------------------------

class SvRecord:
    __slots__ = ("fields", "chrom", "pos", "ref", "alt", "svlen", "svtype", "svread",
                 "names", "gt", "refread", "altread", "reads", "index")

    def __init__(self, fields):
        self.fields = fields
        self.chrom = fields[0]
        self.pos = int(fields[1])
        self.ref = fields[3]
        self.alt = fields[4]
        self.reads = None
        self.index = -1       # position in contig-major order (set by join_support_reads)


def _first_with(info: list[str], needles) -> str | None:
    for item in info:
        for nd in needles:
            if nd in item:
                return item
    return None


def _int_or_zero(txt: str) -> int:
    return 0 if txt == "." else int(txt)


def sv_records(vcf_path: str, include_all_ctgs: bool) -> list[list[SvRecord]]:
    """Per contig, the VCF rows whose CHROM is '<ctg>' or 'chr<ctg>' (:30), with
    SVLEN (:34-36), SVTYPE (:38), support count (:40-47), read names (:48-55) and
    GT / ref-read count (:56-76).  Every "which spelling" decision is made from the
    FIRST record of the contig, as the reference does (re[0], rname[0], gtinfo[0])."""
    names = contig_names(include_all_ctgs, vcf_path[:len(vcf_path) - 24])
    with open(vcf_path, "r") as fh:
        rows = [ln.strip().split() for ln in fh.readlines()]
    out = []
    for ctg in names:
        accept = ("chr" + ctg, ctg)
        recs = [SvRecord(r) for r in rows if r[0] in accept]
        out.append(recs)
        if not recs:
            continue
        infos = [r.fields[7].split(";") for r in recs]
        for r, info in zip(recs, infos):
            item = _first_with(info, ("SVLEN=",))
            if item is None or item == "SVLEN=.":
                item = "SVLEN=0"
            r.svlen = int(item[7:]) if ">" in item else int(item[6:])
            r.svtype = _first_with(info, ("SVTYPE=",))[7:]
        sup_keys = ("SUPPORT=", "SR=", "RE=")
        first = _first_with(infos[0], sup_keys)
        if first is None:
            raise ValueError("no SUPPORT=/RE=/SR= in first record of contig " + ctg)
        cut = 8 if "SUPPORT=" in first else 3
        for r, info in zip(recs, infos):
            r.svread = int(_first_with(info, sup_keys)[cut:])
        name_keys = ("RNAMES=", "READS=")
        first = _first_with(infos[0], name_keys)
        if first is None:
            raise ValueError("no RNAMES=/READS= in first record of contig " + ctg)
        cut = 7 if "RNAMES=" in first else 6
        for r, info in zip(recs, infos):
            r.names = _first_with(info, name_keys)[cut:].split(",")
        samples = [r.fields[9].split(":") for r in recs]
        head = samples[0]
        if len(head) > 4:                       # cuteSV GT:DR:DV:PL:GQ
            mode = "plain"
        elif len(head) >= 3:
            mode = "plain" if head[-1].find(",") == -1 else "ad"   # Sniffles2 / SVIM
        else:
            raise ValueError("unsupported FORMAT/sample column in contig " + ctg)
        for r, smp in zip(recs, samples):
            r.gt = smp[0]
            if mode == "plain":
                r.refread = _int_or_zero(smp[1])
                r.altread = _int_or_zero(smp[2])
            else:
                last = smp[-1]
                cpos = last.find(",")
                # note: with no comma find() is -1, so [:−1] / [0:] -- kept as is
                r.refread = _int_or_zero(last[:cpos])
                r.altread = _int_or_zero(last[cpos + 1:])
    return out


# ---------------------------------------------------------------------------
# the join (sv_phasing_fn.py:36-68)
# ---------------------------------------------------------------------------

def join_support_reads(per_contig: list[list[SvRecord]], tables: list[dict]) -> list[SvRecord]:
    """Each support-read name becomes (name, hap, ps, pc) when the contig's table
    has it, else (name,) (:46-48).  Returns the records flattened in contig order
    with svlen made absolute (:62)."""
    flat = []
    for recs, table in zip(per_contig, tables):
        for r in recs:
            joined = []
            for nm in r.names:
                tag = table.get(nm)
                joined.append((nm,) if tag is None else (nm, tag[0], tag[1], tag[2]))
            r.reads = joined
            r.svlen = abs(r.svlen)
            r.index = len(flat)
            flat.append(r)
    return flat


# ---------------------------------------------------------------------------
# per-SV statistics (sv_phasing_fn.py:70-140) and decision tree (:142-183)
# ---------------------------------------------------------------------------

def nearest_phase_set(sorted_ps, pos):
    """:107-111 -- nearest value to `pos`; an exact tie goes to the upper one."""
    n = len(sorted_ps)
    i = bisect_left(sorted_ps, pos)          # numpy.searchsorted(side='left')
    below = max(i - 1, 0)
    above = min(i, n - 1)
    if abs(pos - sorted_ps[below]) < abs(pos - sorted_ps[above]):
        return sorted_ps[below]
    return sorted_ps[above]


def phase_features(reads, pos, svread, refread, ps_num, oneps, ps_sr=2) -> dict:
    h1 = h2 = t1 = t2 = allhap = hap0 = 0
    ps_pick = 0
    if ps_num == 1:                                             # :74-84
        for rd in reads:
            if len(rd) > 1 and rd[3] <= PC_MAX:
                ps_pick = rd[2]
                if rd[1] == 1:
                    h1 += 1
                    t1 += rd[3]
                elif rd[1] == 2:
                    h2 += 1
                    t2 += rd[3]
        allhap = h1 + h2
    elif ps_num == 2:                                           # :85-105
        stats = {}            # ps -> [tot, n1, n2, sc1, sc2]; insertion ordered
        for rd in reads:
            if len(rd) > 1 and rd[3] <= PC_MAX:
                allhap += 1
                if rd[2] in oneps:
                    st = stats.get(rd[2])
                    if st is None:
                        st = stats[rd[2]] = [0, 0, 0, 0, 0]
                    if rd[1] not in (1, 2):
                        raise KeyError(rd[1])                   # :96
                    st[rd[1]] += 1
                    st[rd[1] + 2] += rd[3]
                    st[0] += 1
        best = 0
        for ps, st in stats.items():                            # first seen wins ties (:101)
            if st[0] > best:
                best = st[0]
                h1, h2, t1, t2, ps_pick = st[1], st[2], st[3], st[4], ps
                hap0 = allhap - h1 - h2
    if ps_num == 0 or (h1 == 0 and h2 == 0):                    # :106-111
        ps_pick = nearest_phase_set(sorted(oneps), pos)
    n = len(reads)
    a1 = t1 / h1 if h1 > 0 else 0                               # :113-116
    a2 = t2 / h2 if h2 > 0 else 0
    lo, hi = min(t1, t2), max(t1, t2)
    f = {
        "hap1": h1, "hap2": h2, "hap0": hap0, "allhap": allhap, "ps": ps_pick,
        "hap1_totsc": t1, "hap2_totsc": t2, "hap1_avgsc": a1, "hap2_avgsc": a2,
        "hapread_ratio": allhap / n,                            # :112
        "tothap": (1 if h1 >= ps_sr else 0) + (2 if h2 >= ps_sr else 0),
        "nohap": n - allhap,
        "hap_diff": abs(h1 - h2),
        "sv_ratio": svread / (svread + refread),                # :123
        "totsc_ratio": hi / lo if lo > 0 else 0,                # :124-125
        "onehap_totsc": hi if lo == 0 else 0,                   # :126-127
        "hap_avgsc_diff": abs(a2 - a1),                         # :132
        "hap_totsc_diff": abs(t2 - t1),
        "ref_num": refread, "sv_num": svread, "allsv": n,
        "hap_ratio": max(h1, h2) / max(min(h1, h2), 1),
        "totsc": t1 + t2,
    }
    amin, amax = min(a1, a2), max(a1, a2)
    f["avgsc_ratio"] = amax / amin if amin > 0 else 0
    f["onehap_avgsc"] = amax if amin == 0 else 0
    f["totsc_ratio2"] = hi / f["totsc"] if f["totsc"] > 0 else 0
    return f


def decide(f: dict, ps_num: int) -> int:
    """T1-T5 thresholds -> 0 drop, 1 '1|0', 2 '0|1', 3 '1|1' (:144-183)."""
    ratio = f["sv_ratio"]
    if ps_num == 0:
        return 3 if (ratio == 1 and f["sv_num"] >= 4) else 0
    if ps_num == 2:
        if ratio >= 0.72:
            if f["hap_avgsc_diff"] <= 1369.50:
                return 3 if f["sv_num"] >= 3 else 0
            return 3 if f["hap0"] >= 6 else 0
        return 0
    # ps_num == 1; the :157-158 test only ever re-assigns 0 and is overwritten below
    pred = 0
    if f["onehap_totsc"] != 0:
        agree = (f["hapread_ratio"] <= 0.75 and f["hap_avgsc_diff"] <= 2400) or f["hapread_ratio"] > 0.75
        if ratio <= 0.24:
            pred = 0
        elif ratio <= 0.9:
            if agree:
                pred = 1 if f["hap1_avgsc"] > 0 else 2
        elif agree:
            pred = 3
    else:
        stronger = 1 if f["hap1_totsc"] > f["hap2_totsc"] else 2
        if ratio <= 0.3:
            pred = 0
        elif ratio <= 0.45:
            pred = 0 if f["ref_num"] > 10 else stronger
        elif ratio <= 0.75:
            pred = 3 if f["totsc_ratio"] <= 9.72 else stronger
        else:
            pred = 3
    return pred


def predict(reads, pos, svread, refread, ps_num, oneps):
    f = phase_features(reads, pos, svread, refread, ps_num, oneps)
    return decide(f, ps_num), f["ps"], f


# ---------------------------------------------------------------------------
# orchestration (sv_phasing_fn.py:185-230)
# ---------------------------------------------------------------------------

def phase_records(flat: list[SvRecord], names: list[str], svlen_thres: int, suppread_thres: int,
                  trace: list | None = None) -> list[dict]:
    """Filter (:189-190), classify by number of distinct PS over ALL joined reads
    (:192-194), collect the per-contig one-PS set from class-1 records' first read
    with pc<=8100 (:195-203), predict contig by contig in class order 0,1,2 skipping
    contigs whose one-PS set is empty (:206-212), stable sort by (chrom, pos) (:229)."""
    kept = [r for r in flat if r.svlen >= svlen_thres and r.svread >= suppread_thres and r.gt != "./."]
    by_class = {0: [], 1: [], 2: []}
    for r in kept:
        distinct = len({rd[2] for rd in r.reads if len(rd) > 1})
        by_class[min(distinct, 2)].append(r)
    oneps = []
    for ctg in names:
        accept = ("chr" + ctg, ctg)
        s = set()
        for r in by_class[1]:
            if r.chrom in accept:
                for rd in r.reads:
                    if len(rd) > 1 and rd[3] <= PC_MAX:
                        s.add(rd[2])
                        break
        oneps.append(s)
    rows = []
    for ci, ctg in enumerate(names):
        accept = ("chr" + ctg, ctg)
        if not oneps[ci]:
            continue
        for ps_num in (0, 1, 2):
            for r in by_class[ps_num]:
                if r.chrom not in accept:
                    continue
                pred, ps, f = predict(r.reads, r.pos, r.svread, r.refread, ps_num, oneps[ci])
                if trace is not None:
                    trace.append((ci, ps_num, r, pred, f))
                if pred == 0:
                    continue
                rows.append({
                    "ps": ps, "hp": GT_TEXT[pred], "chrom": r.chrom, "pos": r.pos,
                    "svlen": r.svlen if r.svtype in ("INS", "DUP") else -r.svlen,
                    "svtype": r.svtype, "ref": r.ref, "alt": r.alt,
                })
    rows.sort(key=lambda d: (d["chrom"], d["pos"]))
    return rows


def generate_phased_callset(vcf_path, sam_home, svlen_thres, suppread_thres, thread, include_all_ctgs,
                            trace: list | None = None) -> list[dict]:
    """Same signature and return value as the reference's function of this name."""
    tables = haplotag_tables(sam_home, thread, include_all_ctgs)
    per_contig = sv_records(vcf_path, include_all_ctgs)
    flat = join_support_reads(per_contig, tables)
    names = contig_names(include_all_ctgs, vcf_path[:len(vcf_path) - 24])
    return phase_records(flat, names, svlen_thres, suppread_thres, trace)


# ---------------------------------------------------------------------------
# output text (write_file.py:6-44)
# ---------------------------------------------------------------------------

_HEADER_FIXED = (
    "##fileformat=VCFv4.2\n"
    "##source=Duet\n"
    '##ALT=<ID=INS,Description="Insertion of novel sequence relative to the reference">\n'
    '##ALT=<ID=DEL,Description="Deletion relative to the reference">\n'
    '##FILTER=<ID=PASS,Description="SV calls passed phasing criterion">\n'
    '##INFO=<ID=SVLEN,Number=1,Type=Integer,Description="Estimated length of the variant">\n'
    '##FORMAT=<ID=HP,Number=1,Type=String,Description="Haplotype of the SV call">\n'
    '##FORMAT=<ID=PS,Number=1,Type=String,Description="Phase set which the SV call belongs to">\n'
)


def header_text(vcf_path: str, include_all_ctgs: bool) -> str:
    names = contig_names(include_all_ctgs, vcf_path[:len(vcf_path) - 24])
    with open(vcf_path, "r") as fh:
        firsts = [ln.strip().split() for ln in fh.readlines()]
    txt = _HEADER_FIXED
    if include_all_ctgs:
        for row in firsts:
            if "##contig=<ID=" in row[0]:
                txt += row[0] + "\n"
    else:
        for ctg in names[:24]:
            for row in firsts:
                if ("##contig=<ID=chr" + ctg + ",") in row[0] or ("##contig=<ID=" + ctg + ",") in row[0]:
                    txt += row[0] + "\n"
    return txt + "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tVALUE\n"


def rows_text(rows: list[dict]) -> str:
    out = []
    for i, c in enumerate(rows, 1):
        out.append("%s\t%s\tDuet.%d\t%s\t%s\t.\tPASS\tSVLEN=%s;SVTYPE=<%s>\tHP:PS\t%s:%s\n" % (
            c["chrom"], c["pos"], i, c["ref"], c["alt"], c["svlen"], c["svtype"], c["hp"], c["ps"]))
    return "".join(out)


def sv_phasing(home, svlen_thres, suppread_thres, thread, include_all_ctgs) -> None:
    """Whole stage: writes <home>/phased_sv.vcf (sv_phasing.py:8-19)."""
    vcf = home + "/sv_calling/variants.vcf"
    rows = generate_phased_callset(vcf, home + "/snp_phasing/", svlen_thres, suppread_thres,
                                   thread, include_all_ctgs)
    with open(home + "/phased_sv.vcf", "w") as fh:
        fh.write(header_text(vcf, include_all_ctgs))
        fh.write(rows_text(rows))
