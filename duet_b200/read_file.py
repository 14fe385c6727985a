"""SV-VCF reader: the host-side decode that replaces the reference's
/root/reference/src/duet/read_file.py (same function names, same argument meaning, same
accepted inputs) but returns COLUMNS for the device instead of nested lists.

Reference behaviour kept (file:line of the reference):
  * contig list = 1..22,X,Y or `tabix --list-chroms` of the pileup VCF          read_file.py:6-16
  * every line is stripped and whitespace-split, header lines included          :18-23
  * a record belongs to contig c when CHROM is 'c' or 'chr'+c                   :30
  * SVLEN: first INFO item containing 'SVLEN=', missing or 'SVLEN=.' -> 0,
    'SVLEN=>N' handled                                                          :34-36
  * SVTYPE: first INFO item containing 'SVTYPE='                                :38
  * support count: first INFO item containing SUPPORT= / SR= / RE=; the prefix
    length (8 or 3) is decided by the FIRST record of the contig                :40-47
  * read names: first item containing RNAMES= / READS=, prefix from the first
    record (7 or 6), split on ','                                               :48-55
  * GT / counts from the sample column, layout decided by the first record      :56-76
    (cuteSV GT:DR:DV:PL:GQ, Sniffles2 GT:GQ:DR:DV -> column [15] is GQ, SVIM GT:DP:AD)
The reference scans all lines once per contig (24 passes) and splits every INFO string into its
items; this reader makes one pass and locates the four items it needs with str.find.
Inputs the reference would silently mis-index (first record of a contig without a support or
read-name item, sample column with fewer than 3 fields) raise ValueError here.
"""
from __future__ import annotations

import shlex
import subprocess
from dataclasses import dataclass, field

import numpy as np

I32_MIN, I32_MAX = -(1 << 31), (1 << 31) - 1


def init_chrom_list(include_all_ctgs, home):
    if not include_all_ctgs:
        return [str(i) for i in range(1, 23)] + ["X", "Y"]
    pileup_vcf_path = home + "/snp_calling/pileup.vcf.gz"
    out = subprocess.check_output(shlex.split("tabix --list-chroms " + pileup_vcf_path))
    return out.decode("ascii").split("\n")[:-1]


def read_file(vcf_path):
    with open(vcf_path, "r") as fh:
        return [ln.split() for ln in fh.readlines()]          # == ln.strip().split() (:20-21)


@dataclass
class ContigSvs:
    """Decoded SV records of one contig, VCF order (one shard's SV side)."""
    chrom: list = field(default_factory=list)     # CHROM string per record
    pos: list = field(default_factory=list)
    ref: list = field(default_factory=list)
    alt: list = field(default_factory=list)
    svlen: list = field(default_factory=list)     # signed, as parsed (:36)
    svtype: list = field(default_factory=list)
    svread: list = field(default_factory=list)
    names_csv: list = field(default_factory=list) # the RNAMES / READS value per record, unsplit ("a,b,c")
    gt: list = field(default_factory=list)
    refread: list = field(default_factory=list)   # column [15]
    altread: list = field(default_factory=list)   # column [16]

    def __len__(self):
        return len(self.pos)

    @property
    def names(self):
        """list[list[str]] -- what the reference keeps in column [13] (:48-55).  The device path hashes the
        unsplit strings natively (duet_hash_name_lists) and never builds these."""
        return [s.split(",") for s in self.names_csv]


def _first_with(info: str, needles):
    """The first ';'-separated item of INFO that contains any of `needles` -- what the reference finds by
    scanning `info.split(';')` item by item (:34-55) -- located with str.find instead: the earliest
    occurrence of a needle lies in exactly that item (no needle contains ';')."""
    at = -1
    for nd in needles:
        k = info.find(nd)
        if k >= 0 and (at < 0 or k < at):
            at = k
    if at < 0:
        return None
    a = info.rfind(";", 0, at) + 1
    b = info.find(";", at)
    return info[a:] if b < 0 else info[a:b]


def _count(txt):
    return 0 if txt == "." else int(txt)


def _i32(v, what):
    if not (I32_MIN <= v <= I32_MAX):
        raise OverflowError(f"{what}={v} does not fit the device's int32 column")
    return v


def parse_vcf(vcf_file, include_all_ctgs):
    """-> list[ContigSvs], one per entry of init_chrom_list (empty when the contig has no record)."""
    chrom_list = init_chrom_list(include_all_ctgs, vcf_file[:len(vcf_file) - 24])
    accept: dict[str, list[int]] = {}
    for ch, c in enumerate(chrom_list):
        for nm in ("chr" + c, c):
            lst = accept.setdefault(nm, [])
            if ch not in lst:
                lst.append(ch)
    buckets: list[list[list[str]]] = [[] for _ in chrom_list]
    for row in read_file(vcf_file):
        for ch in accept.get(row[0], ()):          # row[0] on an empty line raises IndexError, as :30 does
            buckets[ch].append(row)
    out = []
    for ch, rows in enumerate(buckets):
        cs = ContigSvs()
        out.append(cs)
        if not rows:
            continue
        sup_keys, name_keys = ("SUPPORT=", "SR=", "RE="), ("RNAMES=", "READS=")
        first_sup, first_nm = _first_with(rows[0][7], sup_keys), _first_with(rows[0][7], name_keys)
        if first_sup is None or first_nm is None:
            raise ValueError(f"contig {chrom_list[ch]}: first record lacks a SUPPORT=/RE=/SR= or RNAMES=/READS= "
                             "INFO item (the reference mis-indexes its columns on such input)")
        sup_cut = 8 if "SUPPORT=" in first_sup else 3
        nm_cut = 7 if "RNAMES=" in first_nm else 6
        head = rows[0][9].split(":")
        if len(head) > 4:
            ad_mode = False
        elif len(head) >= 3:
            ad_mode = head[-1].find(",") != -1
        else:
            raise ValueError(f"contig {chrom_list[ch]}: sample column '{rows[0][9]}' has fewer than 3 fields")
        svlen_key, svtype_key = ("SVLEN=",), ("SVTYPE=",)
        for r in rows:
            info = r[7]
            item = _first_with(info, svlen_key)
            if item is None or item == "SVLEN=.":
                item = "SVLEN=0"
            cs.svlen.append(int(item[7:]) if ">" in item else int(item[6:]))
            cs.svtype.append(_first_with(info, svtype_key)[7:])
            cs.svread.append(_i32(int(_first_with(info, sup_keys)[sup_cut:]), "support"))
            cs.names_csv.append(_first_with(info, name_keys)[nm_cut:])
            smp = r[9].split(":")
            cs.gt.append(smp[0])
            if ad_mode:
                last = smp[-1]
                k = last.find(",")
                cs.refread.append(_i32(_count(last[:k]), "refread"))
                cs.altread.append(_count(last[k + 1:]))
            else:
                cs.refread.append(_i32(_count(smp[1]), "refread"))
                cs.altread.append(_count(smp[2]))
            cs.chrom.append(r[0])
            cs.pos.append(_i32(int(r[1]), "pos"))
            cs.ref.append(r[3])
            cs.alt.append(r[4])
        for v in cs.svlen:
            _i32(abs(v), "svlen")
    return out


@dataclass
class SvColumns:
    """What the native reader (csrc/vcf_decode.cpp) returns: the records of every listed contig as columns,
    contig-major, VCF order inside a contig, support-read names already hashed."""
    sv_off: np.ndarray        # int64 [n_contigs + 1]
    pos: np.ndarray           # int32 [S]
    svlen: np.ndarray         # int32 [S] signed, as parsed (:36)
    svread: np.ndarray
    refread: np.ndarray
    flags: np.ndarray         # uint8 [S]
    group: np.ndarray         # int32 [S] rank of the CHROM string inside its contig (all zero unless has_groups)
    csr_off: np.ndarray       # int64 [S + 1]
    csr_key: np.ndarray       # uint64 [J]
    csr_chk: np.ndarray       # uint32 [J]
    text: object              # the VCF bytes the spans point into
    str_span: np.ndarray      # int64 [S, 4, 2]: CHROM, REF, ALT, SVTYPE
    has_groups: bool = False


def decode_sv_vcf(vcf_file, include_all_ctgs, threads: int = 1, arena=None):
    """One multi-threaded native pass over the SV VCF -> SvColumns, or None when the file is outside what
    the native reader reproduces exactly (then `parse_vcf`, the general reader, decides -- and raises what
    the reference raises).  `arena(n_svs, n_joins)` supplies the column arrays by name (engine.input_arena:
    views into one page-locked allocation); default: plain numpy arrays."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    chrom_list = init_chrom_list(include_all_ctgs, vcf_file[:len(vcf_file) - 24])
    if any("\n" in c for c in chrom_list) or not chrom_list:
        return None
    with open(vcf_file, "rb") as fh:
        text = fh.read()
    contigs = "\n".join(chrom_list).encode("ascii", "replace")
    job, n_svs, n_joins = C.c_void_p(), C.c_int64(), C.c_int64()
    buf = C.c_char_p(text) if text else C.c_char_p(b"\0")
    rc = lib.duet_decode_sv_vcf(buf, len(text), contigs, len(contigs), max(1, int(threads)), C.byref(job), C.byref(n_svs),
                                C.byref(n_joins))
    if rc != _lib.DUET_OK:
        return None
    S, J, nc = n_svs.value, n_joins.value, len(chrom_list)
    if arena is not None:
        a = arena(S, J)
    else:
        a = {"sv_pos": np.empty(S, np.int32), "sv_svlen": np.empty(S, np.int32), "sv_svread": np.empty(S, np.int32),
             "sv_refread": np.empty(S, np.int32), "sv_flags": np.empty(S, np.uint8), "sv_group": np.empty(S, np.int32),
             "csr_off": np.empty(S + 1, np.int64), "csr_key": np.empty(J, np.uint64), "csr_chk": np.empty(J, np.uint32)}
    cols = SvColumns(np.empty(nc + 1, np.int64), a["sv_pos"], a["sv_svlen"], a["sv_svread"], a["sv_refread"], a["sv_flags"],
                     a["sv_group"], a["csr_off"], a["csr_key"], a["csr_chk"], text, np.empty((S, 4, 2), np.int64))
    has_groups = C.c_int32()
    lib.duet_svs_take(job, cols.sv_off.ctypes.data, cols.pos.ctypes.data, cols.svlen.ctypes.data, cols.svread.ctypes.data,
                      cols.refread.ctypes.data, cols.flags.ctypes.data, cols.group.ctypes.data, C.byref(has_groups),
                      cols.csr_off.ctypes.data, cols.csr_key.ctypes.data, cols.csr_chk.ctypes.data, cols.str_span.ctypes.data)
    cols.has_groups = bool(has_groups.value)
    return cols
