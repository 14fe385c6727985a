"""Seeded synthetic inputs for the sv_phasing hot path.

One generator emits BOTH views of the same data:

  * the on-disk artefacts the reference consumes -- ``<home>/snp_phasing/<ctg>.bam``
    written as SAM *text* (the reference pipes ``samtools view`` text,
    /root/reference/src/duet/sv_phasing_fn.py:25) and
    ``<home>/sv_calling/variants.vcf`` in the cuteSV / Sniffles2 / SVIM dialects
    (/root/reference/src/duet/read_file.py:40-76);
  * the per-contig arrays (`SynthSample`) from which ``duet_b200.columnar``
    builds the device batch without going through text (bench at WGS scale).

Read names are 36-character UUID-like strings derived from an integer read id,
so the text view and the hashed view always agree.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

GRCH37 = {
    "1": 249250621, "2": 243199373, "3": 198022430, "4": 191154276, "5": 180915260,
    "6": 171115067, "7": 159138663, "8": 146364022, "9": 141213431, "10": 135534747,
    "11": 135006516, "12": 133851895, "13": 115169878, "14": 107349540, "15": 102531392,
    "16": 90354753, "17": 81195210, "18": 78077248, "19": 59128983, "20": 63025520,
    "21": 48129895, "22": 51304566, "X": 155270560, "Y": 59373566,
}
CHROM_LIST = [str(i) for i in range(1, 23)] + ["X", "Y"]

SVTYPES = ["DEL", "INS", "DUP", "INV", "BND"]
GTS = ["0/1", "1/1", "0/0", "./."]

_HEX = np.frombuffer(b"0123456789abcdef", dtype=np.uint8)


def _splitmix(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def names_from_ids(ids: np.ndarray) -> np.ndarray:
    """(N,) int read ids -> (N, 36) uint8 UUID-like ASCII names, injective in id."""
    ids = np.asarray(ids, dtype=np.uint64)
    a = _splitmix(ids)  # scrambled half, for realistic looking names
    b = ids             # the id itself keeps the mapping injective
    nib = np.empty((ids.shape[0], 32), dtype=np.uint8)
    for i in range(16):
        nib[:, i] = (a >> np.uint64(60 - 4 * i)) & np.uint64(15)
        nib[:, 16 + i] = (b >> np.uint64(60 - 4 * i)) & np.uint64(15)
    hx = _HEX[nib]
    out = np.empty((ids.shape[0], 36), dtype=np.uint8)
    out[:, 0:8] = hx[:, 0:8]
    out[:, 8] = ord("-")
    out[:, 9:13] = hx[:, 8:12]
    out[:, 13] = ord("-")
    out[:, 14:18] = hx[:, 12:16]
    out[:, 18] = ord("-")
    out[:, 19:23] = hx[:, 16:20]
    out[:, 23] = ord("-")
    out[:, 24:36] = hx[:, 20:32]
    return out


def name_strings(ids: np.ndarray) -> list[str]:
    arr = names_from_ids(ids)
    return [r.tobytes().decode("ascii") for r in arr]


@dataclass
class SynthContig:
    name: str                      # contig name as in chrom_list ('1', 'X', ...)
    # haplotagged BAM rows in file order
    row_id: np.ndarray             # int64 read id
    row_pos: np.ndarray            # int32 alignment start
    row_tagged: np.ndarray         # bool: carries HP/PC/PS
    row_hp: np.ndarray             # uint8
    row_ps: np.ndarray             # int32
    row_pc: np.ndarray             # int32
    # SV records in VCF order
    sv_pos: np.ndarray             # int32
    sv_len: np.ndarray             # int32 signed SVLEN as written (0 => no SVLEN field)
    sv_type: np.ndarray            # int8 index into SVTYPES
    sv_gt: np.ndarray              # int8 index into GTS
    sv_refread: np.ndarray         # int32
    sv_svread: np.ndarray          # int32 (RE / SUPPORT value)
    sup_off: np.ndarray            # int64 CSR offsets into sup_id
    sup_id: np.ndarray             # int64 read ids of support reads


@dataclass
class SynthSample:
    contigs: list[SynthContig]
    chr_prefix: bool = False       # write 'chr1' instead of '1'
    seed: int = 0
    meta: dict = field(default_factory=dict)

    @property
    def n_rows(self) -> int:
        return int(sum(c.row_id.shape[0] for c in self.contigs))

    @property
    def n_tagged(self) -> int:
        return int(sum(int(c.row_tagged.sum()) for c in self.contigs))

    @property
    def n_svs(self) -> int:
        return int(sum(c.sv_pos.shape[0] for c in self.contigs))

    @property
    def n_joins(self) -> int:
        return int(sum(c.sup_id.shape[0] for c in self.contigs))


def _support_count(rng, n, dense):
    if not dense:
        return rng.integers(2, 31, size=n)
    # mean ~60 with a heavy tail to 2000
    k = np.minimum(2000, np.maximum(2, (rng.lognormal(mean=3.55, sigma=0.95, size=n)).astype(np.int64)))
    return k


def make_contig(rng: np.random.Generator, name: str, length: int, n_reads: int, n_svs: int,
                id_base: int, *, dense: bool = False, tagged_frac: float = 0.8,
                dup_frac: float = 0.02, empty_oneps: bool = False,
                block_mean: float = 500_000.0, shuffle_vcf: bool = False) -> SynthContig:
    n_reads = max(int(n_reads), 4)
    pos = np.sort(rng.integers(1, length, size=n_reads)).astype(np.int32)
    ids = id_base + np.arange(n_reads, dtype=np.int64)
    # phase blocks: start positions, Exp(block_mean) lengths
    n_blk = int(length / block_mean * 1.5) + 8
    starts = np.cumsum(rng.exponential(block_mean, size=n_blk)).astype(np.int64)
    starts = np.concatenate([[0], starts[starts < length]])
    blk_ps = (starts + 1 + rng.integers(0, 1000, size=starts.shape[0])).astype(np.int32)
    ps = blk_ps[np.searchsorted(starts, pos, side="right") - 1]
    # ~10 % of phase blocks are "untaggable" (no het SNPs nearby): SVs there have no
    # haplotagged support read at all (class 0 in sv_phasing_fn.py:194)
    blk_dead = rng.random(starts.shape[0]) < 0.1
    p_tag = np.where(blk_dead[np.searchsorted(starts, pos, side="right") - 1], 0.02,
                     min(1.0, tagged_frac / 0.902))
    tagged = rng.random(n_reads) < p_tag
    hp = rng.integers(1, 3, size=n_reads).astype(np.uint8)
    pc = np.minimum(rng.exponential(1500.0, size=n_reads), 20000.0).astype(np.int32)
    pc[rng.random(n_reads) < 0.01] = 0
    pc[rng.random(n_reads) < 0.002] = 8100
    pc[rng.random(n_reads) < 0.002] = 8101
    if empty_oneps:
        pc[:] = 9000
    # duplicate QNAME rows (supplementary alignments) carrying a different tag
    n_dup = int(n_reads * dup_frac)
    if n_dup:
        src = rng.integers(0, n_reads, size=n_dup)
        d_pos = rng.integers(1, length, size=n_dup).astype(np.int32)
        d_ps = blk_ps[np.searchsorted(starts, d_pos, side="right") - 1]
        near = rng.random(n_dup) < 0.5      # half stay in the same phase block
        d_ps = np.where(near, ps[src], d_ps)
        d_pos = np.where(near, np.minimum(pos[src] + rng.integers(0, 50, size=n_dup), length - 1), d_pos)
        d_hp = (3 - hp[src]).astype(np.uint8)
        d_pc = np.minimum(rng.exponential(1500.0, size=n_dup), 20000.0).astype(np.int32)
        if empty_oneps:
            d_pc[:] = 9000
        d_tag = rng.random(n_dup) < 0.9
        ids = np.concatenate([ids, ids[src]])
        pos = np.concatenate([pos, d_pos.astype(np.int32)])
        ps = np.concatenate([ps, d_ps.astype(np.int32)])
        hp = np.concatenate([hp, d_hp])
        pc = np.concatenate([pc, d_pc])
        tagged = np.concatenate([tagged, d_tag])
        order = np.argsort(pos, kind="stable")
        ids, pos, ps, hp, pc, tagged = ids[order], pos[order], ps[order], hp[order], pc[order], tagged[order]
    n_rows = ids.shape[0]

    # SV records
    n_svs = int(n_svs)
    sv_pos = np.sort(rng.integers(1000, max(length - 1000, 2000), size=n_svs)).astype(np.int32)
    k = _support_count(rng, n_svs, dense)
    sv_type = rng.choice(len(SVTYPES), size=n_svs, p=[0.44, 0.44, 0.05, 0.05, 0.02]).astype(np.int8)
    sv_abs = rng.integers(30, 5001, size=n_svs).astype(np.int32)
    sv_len = np.where(sv_type == 0, -sv_abs, sv_abs).astype(np.int32)
    sv_len[sv_type == 4] = 0                      # BND: no SVLEN field
    sv_gt = rng.choice(len(GTS), size=n_svs, p=[0.55, 0.3, 0.1, 0.05]).astype(np.int8)
    het = rng.random(n_svs) < 0.6
    # het SVs see about as many reference reads as support reads, hom SVs few or none
    ref_het = (k * rng.uniform(0.4, 1.6, size=n_svs)).astype(np.int64)
    ref_hom = np.where(rng.random(n_svs) < 0.6, 0, rng.integers(0, 4, size=n_svs))
    sv_refread = np.where(het, ref_het, ref_hom)
    noisy = rng.random(n_svs) < 0.15
    sv_refread = np.where(noisy, rng.integers(0, 41, size=n_svs), sv_refread).astype(np.int32)
    pref = rng.integers(1, 3, size=n_svs)
    centre = np.searchsorted(pos, sv_pos)
    sup_lists = []
    absent_base = id_base + (1 << 40)
    for i in range(n_svs):
        ki = int(k[i])
        lo = max(0, int(centre[i]) - ki)
        hi = min(n_rows, lo + 2 * ki)
        lo = max(0, hi - 2 * ki)
        cand = np.arange(lo, hi)
        if cand.shape[0] <= ki:
            pick = cand
        else:
            if het[i]:
                w = np.where(tagged[cand], np.where(hp[cand] == pref[i], 0.97, 0.03), 0.5)
            else:
                w = np.full(cand.shape[0], 0.5)
            keyv = rng.random(cand.shape[0]) ** (1.0 / w)
            pick = cand[np.argpartition(-keyv, ki - 1)[:ki]]
            pick = pick[rng.permutation(pick.shape[0])]
        lst = ids[pick]
        r = rng.random()
        if r < 0.05:      # names absent from this contig's BAM
            extra = absent_base + rng.integers(0, 1 << 30, size=int(rng.integers(1, 3)))
            lst = np.concatenate([lst, extra])
            lst = lst[rng.permutation(lst.shape[0])]
        elif r < 0.08 and lst.shape[0] > 0:   # the same read listed twice
            lst = np.concatenate([lst, lst[:1]])
        sup_lists.append(lst.astype(np.int64))
    sv_svread = np.array([l.shape[0] for l in sup_lists], dtype=np.int32)
    # a few records whose RE/SUPPORT value differs from the list length
    odd = rng.random(n_svs) < 0.02
    sv_svread = np.where(odd, sv_svread + rng.integers(1, 4, size=n_svs), sv_svread).astype(np.int32)
    if shuffle_vcf and n_svs > 1:
        perm = rng.permutation(n_svs)
        sv_pos, sv_len, sv_type, sv_gt, sv_refread, sv_svread = (a[perm] for a in
            (sv_pos, sv_len, sv_type, sv_gt, sv_refread, sv_svread))
        sup_lists = [sup_lists[j] for j in perm]
    sup_off = np.zeros(n_svs + 1, dtype=np.int64)
    if n_svs:
        sup_off[1:] = np.cumsum([l.shape[0] for l in sup_lists])
    sup_id = np.concatenate(sup_lists) if sup_lists else np.zeros(0, dtype=np.int64)
    return SynthContig(name, ids, pos, tagged, hp, ps.astype(np.int32), pc.astype(np.int32),
                       sv_pos, sv_len, sv_type, sv_gt, sv_refread, sv_svread, sup_off, sup_id)


def make_sample(seed: int = 0, *, contigs=None, n_reads: int = 70_000, n_svs: int = 2_500,
                dense: bool = False, chr_prefix: bool = False, empty_oneps_contig: str | None = "auto",
                shuffle_vcf: bool = False, id_base: int = 0, bp_per_read: float | None = None,
                **kw) -> SynthSample:
    """Reads and SVs are spread over `contigs` proportionally to GRCh37 length.  With
    `bp_per_read` the contigs are shrunk to keep that read density (small test cases
    then look like 30x data: ~700 bp between read starts)."""
    rng = np.random.default_rng(seed)
    contigs = list(contigs) if contigs is not None else ["21"]
    total = float(sum(GRCH37.get(c, 50_000_000) for c in contigs))
    if empty_oneps_contig == "auto":
        empty_oneps_contig = contigs[-1] if len(contigs) > 1 else None
    out = []
    base = id_base
    for c in contigs:
        length = GRCH37.get(c, 50_000_000)
        nr = int(round(n_reads * length / total))
        ns = int(round(n_svs * length / total))
        if bp_per_read is not None:
            length = max(int(nr * bp_per_read), 20_000)
        sc = make_contig(rng, c, length, nr, ns, base, dense=dense,
                         empty_oneps=(c == empty_oneps_contig), shuffle_vcf=shuffle_vcf, **kw)
        out.append(sc)
        base += 1 << 32
    return SynthSample(out, chr_prefix=chr_prefix, seed=seed,
                       meta=dict(n_reads=n_reads, n_svs=n_svs, dense=dense))


# named shapes of BASELINE.json `configs`
def config_c1(seed=0, **kw):
    return make_sample(seed, contigs=["21"], n_reads=70_000, n_svs=2_500, **kw)


def config_c2(seed=0, **kw):
    return make_sample(seed, contigs=CHROM_LIST, n_reads=4_500_000, n_svs=25_000, **kw)


def config_c4(seed=0, **kw):
    return make_sample(seed, contigs=CHROM_LIST, n_reads=9_000_000, n_svs=30_000, dense=True, **kw)


# ----------------------------------------------------------------------------
# text writers (the reference's input formats)
# ----------------------------------------------------------------------------

def write_sam_text(c: SynthContig, path: str) -> None:
    """One haplotagged 'BAM' as SAM text; tagged rows end in HP:i PC:i PS:i in the
    order `whatshap haplotag` appends them (sv_phasing_fn.py:28-29 reads s[-3:])."""
    names = name_strings(c.row_id)
    with open(path, "w") as f:
        w = f.write
        for i, nm in enumerate(names):
            head = f"{nm}\t0\t{c.name}\t{int(c.row_pos[i])}\t60\t4M\t*\t0\t0\tACGT\tIIII\tNM:i:0"
            if c.row_tagged[i]:
                w(f"{head}\tHP:i:{int(c.row_hp[i])}\tPC:i:{int(c.row_pc[i])}\tPS:i:{int(c.row_ps[i])}\n")
            else:
                w(f"{head}\tMD:Z:4\tAS:i:8\n")


def _vcf_line(dialect: str, chrom: str, idx: int, pos: int, svlen: int, svtype: str, gt: str,
              refread: int, svread: int, names: list[str]) -> str:
    rn = ",".join(names)
    lenfield = "" if svtype == "BND" else f"SVLEN={svlen};"
    alt = f"<{svtype}>"
    dv = len(names)
    if dialect == "cutesv":
        info = f"PRECISE;SVTYPE={svtype};{lenfield}END={pos + abs(svlen)};CIPOS=0,0;CILEN=0,0;RE={svread};RNAMES={rn};STRAND=+-"
        fmt, smp = "GT:DR:DV:PL:GQ", f"{gt}:{refread}:{dv}:10,0,10:10"
    elif dialect == "sniffles":
        info = (f"PRECISE;SVTYPE={svtype};{lenfield}END={pos + abs(svlen)};SUPPORT={svread};RNAMES={rn};"
                f"COVERAGE=20,20,20,20,20;STRAND=+-;AF=0.500;STDEV_LEN=1.2;STDEV_POS=0.5")
        fmt, smp = "GT:GQ:DR:DV", f"{gt}:{refread}:{refread + 1}:{dv}"
    elif dialect == "svim":
        info = f"SVTYPE={svtype};END={pos + abs(svlen)};{lenfield}SUPPORT={svread};STD_SPAN=1.0;STD_POS=2.0;READS={rn}"
        fmt, smp = "GT:DP:AD", f"{gt}:{refread + dv}:{refread},{dv}"
    else:
        raise ValueError(dialect)
    return f"{chrom}\t{pos}\t{dialect}.{svtype}.{idx}\tN\t{alt}\t.\tPASS\t{info}\t{fmt}\t{smp}\n"


def write_vcf(sample: SynthSample, path: str, dialect: str = "cutesv") -> None:
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.2\n##source=synthetic\n")
        for c in sample.contigs:
            cn = ("chr" if sample.chr_prefix else "") + c.name
            f.write(f"##contig=<ID={cn},length={GRCH37.get(c.name, 50_000_000)}>\n")
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n")
        idx = 0
        for c in sample.contigs:
            cn = ("chr" if sample.chr_prefix else "") + c.name
            names = name_strings(c.sup_id)
            for i in range(c.sv_pos.shape[0]):
                lst = names[int(c.sup_off[i]):int(c.sup_off[i + 1])]
                f.write(_vcf_line(dialect, cn, idx, int(c.sv_pos[i]), int(c.sv_len[i]),
                                  SVTYPES[int(c.sv_type[i])], GTS[int(c.sv_gt[i])],
                                  int(c.sv_refread[i]), int(c.sv_svread[i]), lst))
                idx += 1


def write_workdir(sample: SynthSample, home: str, dialect: str = "cutesv") -> None:
    """Lay out `<home>/snp_phasing/<ctg>.bam` (SAM text) and `<home>/sv_calling/variants.vcf`."""
    os.makedirs(os.path.join(home, "snp_phasing"), exist_ok=True)
    os.makedirs(os.path.join(home, "sv_calling"), exist_ok=True)
    for c in sample.contigs:
        fn = ("chr" if sample.chr_prefix else "") + c.name + ".bam"
        write_sam_text(c, os.path.join(home, "snp_phasing", fn))
    write_vcf(sample, os.path.join(home, "sv_calling", "variants.vcf"), dialect)


# ----------------------------------------------------------------------------
# SV signatures for kernel set B (BASELINE.json configs[2])
# ----------------------------------------------------------------------------

def make_signatures(seed: int = 0, n: int = 2_000_000, contigs=None, shuffle: bool = True):
    """True events Poisson along each contig, 1-40 signatures per event, start jitter N(0, 50),
    span lognormal(median 300) x (1 +- 0.1); types DEL/INS/INV/DUP_TAN 45/45/5/5 %.
    Returns int32 arrays (contig, type, start, end); insertions use end = start + length."""
    rng = np.random.default_rng(seed)
    contigs = list(contigs) if contigs is not None else CHROM_LIST
    lengths = np.array([GRCH37.get(c, 50_000_000) for c in contigs], np.float64)
    per = rng.integers(1, 41, size=int(n / 20.5) + 64)
    per = per[np.cumsum(per) <= n]
    n_ev = per.shape[0]
    ev_contig = rng.choice(len(contigs), size=n_ev, p=lengths / lengths.sum())
    ev_pos = (rng.random(n_ev) * (lengths[ev_contig] - 20_000) + 10_000).astype(np.int64)
    ev_type = rng.choice(4, size=n_ev, p=[0.45, 0.45, 0.05, 0.05])
    ev_span = np.maximum(30, rng.lognormal(np.log(300.0), 0.8, size=n_ev)).astype(np.int64)
    ev = np.repeat(np.arange(n_ev), per)
    m = ev.shape[0]
    start = ev_pos[ev] + np.rint(rng.normal(0.0, 50.0, size=m)).astype(np.int64)
    span = np.maximum(1, np.rint(ev_span[ev] * (1.0 + rng.uniform(-0.1, 0.1, size=m)))).astype(np.int64)
    contig, typ = ev_contig[ev], ev_type[ev]
    if shuffle:
        p = rng.permutation(m)
        contig, typ, start, span = contig[p], typ[p], start[p], span[p]
    start = np.maximum(start, 0)
    return (contig.astype(np.int32), typ.astype(np.int32), start.astype(np.int32), (start + span).astype(np.int32))
